"""Statistics container with the key layout of the reference (``pySDC/core/hooks.py:8-19`` ``Entry`` namedtuple) and the
filter / sort helpers user scripts rely on (``pySDC/helpers/stats_helper.py:4-111``)."""
from collections import namedtuple

Entry = namedtuple("Entry", ["process", "process_sweeper", "time", "level", "iter", "sweep", "type", "num_restarts"])


def filter_stats(stats, **kwargs):
    out = {}
    for k, v in stats.items():
        if all(getattr(k, name) == want for name, want in kwargs.items() if name != "comm"):
            out[k] = v
    return out


def sort_stats(stats, sortby="time"):
    return sorted(((getattr(k, sortby), v) for k, v in stats.items()), key=lambda kv: kv[0])


def get_sorted(stats, sortby="time", **kwargs):
    return sort_stats(filter_stats(stats, **kwargs), sortby=sortby)


def get_list_of_types(stats):
    return sorted({k.type for k in stats})
