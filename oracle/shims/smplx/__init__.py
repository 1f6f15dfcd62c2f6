"""Minimal stand-in for the third-party `smplx` package (pinned smplx==0.1.26 in the
reference's requirements.txt:7; absent from this image).  It routes `smplx.lbs.lbs`
to the oracle's restatement so that the UNMODIFIED reference python
(head_detector/flame.py:8-10,152-161) can be imported and run here to generate
golden vectors.  TEST INFRASTRUCTURE ONLY - never imported by the product."""
