#!/usr/bin/env python
"""Shared-memory wavefront model of the x / y exchange stages of `k_apply3d_group` (orders 4-6, two-warp groups).

A 64-bit access of a warp is served per half-warp; a half-warp needs as many wavefronts as the largest number of
distinct addresses that fall into one of the 16 eight-byte banks.  The script counts wavefronts per element for the four
access shapes (F1 write / B3 read of the x-lines, F2 / B2 reads and writes of the (qx,dz) role, F3 read / B1 write of the
(qx,qy) columns) and prints the layouts used in `csrc/kernels_apply_warp.cu` next to the ideal count.

    python scripts/bank_model.py
"""


def wavefronts(addrs):
    tot = 0
    for hw in range(4):
        a = {ad for (lane, ad) in addrs if hw * 16 <= lane < hw * 16 + 16}
        if not a:
            continue
        banks = {}
        for ad in a:
            banks[ad % 16] = banks.get(ad % 16, 0) + 1
        tot += max(banks.values())
    return tot


def cost(P, l1, l2, p_addr, r_stride_y, rstr):
    """l1: (lane, dy, dz) of the x-line role; l2: (lane, qx, dz); p_addr(dy, dz): row of an x-line in the P layout"""
    D, Q = P + 1, P + 2
    l3 = [(t, t % Q, t // Q) for t in range(Q * Q)]
    c = {"F1w/B3r": wavefronts([(l, p_addr(dy, dz)) for l, dy, dz in l1]) * (Q * 4),
         "F2r/B2w": wavefronts([(l, qx + p_addr(0, dz)) for l, qx, dz in l2]) * (D * 4) if p_addr(1, 0) != r_stride_y
         else wavefronts([(l, qx + rstr * dz) for l, qx, dz in l2]) * (D * 4),
         "F2w/B2r": wavefronts([(l, qx + rstr * dz) for l, qx, dz in l2]) * (Q * 6),
         "F3r/B1w": wavefronts([(l, qx + r_stride_y * qy) for l, qx, qy in l3]) * (D * 6)}
    return sum(c.values()), c


def ideal(P):
    D, Q = P + 1, P + 2
    h = lambda n: (n + 15) // 16
    return h(D * D) * Q * 4 + h(Q * D) * (D * 4 + Q * 6) + h(Q * Q) * D * 6


def main():
    # p=4: separate P buffers, P(q; line) = q + 9 line, R stride 45, dense lane maps
    P, D, Q = 4, 5, 6
    l1 = [(t, t % D, t // D) for t in range(D * D)]
    l2 = [(t, t // D, t % D) for t in range(Q * D)]
    print("p=4 separate P (PST 9, RSTR 45):", cost(P, l1, l2, lambda dy, dz: 9 * (dy + D * dz), Q, 45), "ideal", ideal(P))
    # p=5: in place, dense maps (RSTR 58) and padded half-warp maps (RSTR 56)
    P, D, Q = 5, 6, 7
    l1 = [(t, t % D, t // D) for t in range(D * D)]
    l2 = [(t, t // D, t % D) for t in range(Q * D)]
    print("p=5 in place, dense maps (RSTR 58):", cost(P, l1, l2, lambda dy, dz: Q * dy + 58 * dz, Q, 58), "ideal", ideal(P))
    l1 = [(16 * h + r, r % D, 2 * h + r // D) for h in range(3) for r in range(12)]
    l2 = [(16 * h + r, r % Q, 2 * h + r // Q) for h in range(3) for r in range(14)]
    print("p=5 in place, padded maps (RSTR 56):", cost(P, l1, l2, lambda dy, dz: Q * dy + 56 * dz, Q, 56))
    # p=6: in place, RSTR 71
    P, D, Q = 6, 7, 8
    l1 = [(t, t // D, t % D) for t in range(D * D)]
    l2 = [(t, t // D, t % D) for t in range(Q * D)]
    print("p=6 in place, dense maps (RSTR 71):", cost(P, l1, l2, lambda dy, dz: Q * dy + 71 * dz, Q, 71), "ideal", ideal(P))
    l1 = [(16 * h + r, 2 * h + r // D, r % D) for h in range(4) for r in range(14) if 2 * h + r // D < D]
    print("p=6 in place, padded x-line map (RSTR 71):", cost(P, l1, l2, lambda dy, dz: Q * dy + 71 * dz, Q, 71))


if __name__ == "__main__":
    main()
