"""numpy emulation of the sweep-pair schedule of csrc/net_kernels.cu::k_rotf (same index arithmetic and step order),
so that every combination of the two sweeps' column ranges can be checked against applying the sweeps one after the
other.  Test infrastructure only."""
import numpy as np


def rot(cy, zi, e):
    """k_rot's rot4: columns (i, i+1) = (zi, cy) -> (new col i+1, new col i)."""
    c, s = e
    return s * zi + c * cy, c * zi - s * cy


def apply_single(Z, m, params):
    """one sweep: rotations on (i, i+1) for i = m-1, m-2, ..., m-cnt (the unfused k_rot)."""
    carry = Z[:, m].copy()
    for t, e in enumerate(params):
        i = m - 1 - t
        out, carry = rot(carry, Z[:, i], e)
        Z[:, i + 1] = out
    Z[:, m - len(params)] = carry


def apply_pair(Z, mA, pa, mB, pb):
    """the fused schedule: sweep B trails sweep A by one column (k_rotf, `if (pair)` branch)."""
    cntA, cntB = len(pa), len(pb)
    lA, lB = mA - cntA, mB - cntB
    itop, iend = max(mA, mB) - 1, min(lA, lB - 1)
    cA = Z[:, itop + 1].copy()
    zi = Z[:, itop].copy()
    if lA <= itop <= mA - 1:
        cB, cA = rot(cA, zi, pa[mA - 1 - itop])
    else:
        cB, cA = cA, zi
    core_lo, core_hi = max(lA, lB - 1), min(mA - 1, mB - 2)

    def flagged(i, cA, cB):
        zi = Z[:, i].copy() if i >= 0 else np.zeros(Z.shape[0])
        if lA <= i <= mA - 1:
            oA, cA = rot(cA, zi, pa[mA - 1 - i])
        else:
            oA, cA = cA, zi
        if lB <= i + 1 <= mB - 1:
            oB, cB = rot(cB, oA, pb[mB - 2 - i])
        else:
            oB, cB = cB, oA
        Z[:, i + 2] = oB
        return cA, cB

    i = itop - 1
    while i >= iend and i > core_hi:
        cA, cB = flagged(i, cA, cB)
        i -= 1
    while i - 1 >= core_lo:                       # two steps per trip, both sweeps active
        z0, z1 = Z[:, i].copy(), Z[:, i - 1].copy()
        oA0, cA = rot(cA, z0, pa[mA - 1 - i])
        oB0, cB = rot(cB, oA0, pb[mB - 2 - i])
        oA1, cA = rot(cA, z1, pa[mA - i])
        oB1, cB = rot(cB, oA1, pb[mB - 1 - i])
        Z[:, i + 2], Z[:, i + 1] = oB0, oB1
        i -= 2
    while i >= iend:
        cA, cB = flagged(i, cA, cB)
        i -= 1
    Z[:, iend + 1] = cB
    if iend >= 0:
        Z[:, iend] = cA


def would_pair(mA, cntA, mB, cntB):
    top, end = max(mA, mB) - 1, min(mA - cntA, mB - cntB - 1)
    return cntA > 0 and cntB > 0 and 20 * (top - end + 1) <= 17 * (cntA + cntB)
