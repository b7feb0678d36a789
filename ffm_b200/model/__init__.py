"""Drop-in replacements for the reference's ``model/ffm_*.py`` classes.

Same constructor, methods and attributes; the stepping is done by the CUDA
rollout kernels through the C ABI.  To run the reference's drivers unmodified, put this package's
parent directory first on ``sys.path`` under the name ``model`` or import
``from ffm_b200.model.ffm_core import FloorFieldModel``.
"""
