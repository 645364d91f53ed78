"""Stub so the reference's cns/benchmark/stop_policy.py imports here (scikit-image is not installed);
only SSIMStopPolicy, which is out of scope, would call into it."""
