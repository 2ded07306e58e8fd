from .bad import BADDescriptor, SparseBAD, extract_descriptors_at_keypoints, extract_descriptors_at_keypoints_subpixel

__all__ = ["BADDescriptor", "SparseBAD", "extract_descriptors_at_keypoints",
           "extract_descriptors_at_keypoints_subpixel"]
