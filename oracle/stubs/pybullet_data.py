"""Empty stand-in (see pybullet.py)."""
def getDataPath():
    return ""
