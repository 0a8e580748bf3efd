"""Drop-in module ``process_image_cpp`` (module name of setup.py:29 / process_image.cpp:369).

``import process_image_cpp; process_image_cpp.process_image_cpp(...)`` keeps working for callers of the
reference extension (spacevoxelviewer.py:233-251); the work runs on the B200 through libpvp_b200.so.
"""
from pixeltovoxelprojector_b200.process_image import process_image_cpp  # noqa: F401

__doc__ = "B200 implementation of the process_image function"
