"""Stand-in: only what the reference touches at import time (cns/utils/visualize.py:16-18)."""


class _Cycle(object):
    def by_key(self):
        return {"color": ["#1f77b4", "#ff7f0e", "#2ca02c", "#d62728", "#9467bd", "#8c564b", "#e377c2", "#7f7f7f",
                          "#bcbd22", "#17becf"]}


rcParams = {"axes.prop_cycle": _Cycle()}
