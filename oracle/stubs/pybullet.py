"""Empty stand-in so reference modules that `import pybullet` at top level import here.
TEST INFRASTRUCTURE ONLY (oracle); nothing on the hot path calls pybullet."""
GUI_SERVER = 0
ER_BULLET_HARDWARE_OPENGL = 0
def __getattr__(name):
    raise AttributeError("pybullet stub: %s is not available" % name)
