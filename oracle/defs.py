"""Enumerations shared by the oracle tiers (kept numerically equal to include/qw_b200.h;
tests/test_abi.py checks that)."""

# QW_Search.py:70-81 -- topology 0 = complete graph, 1 = cycle graph
TOPOLOGY_COMPLETE = 0
TOPOLOGY_CYCLE = 1

# QW_Search.py:104-109 fixes 0 = cerf_3, 1 = lin, 2 = sqrt; the rest continue the list with
# the remaining BCB.py:75-87 schedules and the time-independent comparator (BCB.py:101-135).
SCHEDULES = {
    "cerf_3": 0,
    "lin": 1,
    "sqrt": 2,
    "cbrt": 3,
    "cerf_5": 4,
    "cerf_7": 5,
    "static": 6,
}
SCHEDULE_NAMES = {v: k for k, v in SCHEDULES.items()}

# where gamma sits: on the oracle term (all reference code: BCB.py:93-97) or on the
# Laplacian (thesis Eq. 2.14, p.47 -- no reference code).
GAMMA_ON_ORACLE = 0
GAMMA_ON_LAPLACIAN = 1


def schedule_id(s):
    """Accept the three encodings of the reference (SURVEY.md section 5 'Config')."""
    if isinstance(s, str):
        key = {"independent": "static"}.get(s, s)
        return SCHEDULES[key]
    return int(s)


def topology_id(t):
    if isinstance(t, str):
        return {"complete": TOPOLOGY_COMPLETE, "circular": TOPOLOGY_CYCLE,
                "cycle": TOPOLOGY_CYCLE}[t]
    return int(t)


def gamma_on_id(g):
    if isinstance(g, str):
        return {"oracle": GAMMA_ON_ORACLE, "laplacian": GAMMA_ON_LAPLACIAN}[g]
    return int(g)


def default_target(n):
    """BCB.py:96-97 -- the marked vertex is always int((N-1)/2)."""
    return int((n - 1) / 2)
