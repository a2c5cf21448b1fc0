"""Minimal stand-in for the `timm` distribution (pinned timm==1.0.3 in the reference's pyproject.toml:24), which is not
installed in this image.  Only what videollama2/model/projector.py:22-23 imports is provided.  TEST INFRASTRUCTURE."""
__version__ = "1.0.3-shim"
