"""TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's Shi-Tomasi -> NMS -> top-k -> BAD -> Sinkhorn path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (onnx_image_processing_b200) never does.
"""
