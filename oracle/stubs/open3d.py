"""Empty stand-in so `cns/utils/perception.py` imports; CameraIntrinsic needs none of it."""
