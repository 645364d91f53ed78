"""Empty stand-in (reference sim modules import matplotlib.pyplot at top level)."""
