"""What the grouping kernels face on a benchmark frame: peaks per part, candidate pairs with a positive score per limb,
connections made, and how many parallel rounds the matching needs (emulated on the host from the device's pair-score matrix).
usage: python tools/debug_group_stats.py [C2|C3]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import isl_b200  # noqa: E402
from isl_b200 import synth, tables  # noqa: E402


def rounds_needed(m):
    """Rounds of csrc/group.cu match_kernel on the score matrix m (nA x nB, -1 = not a candidate)."""
    freeA = np.ones(m.shape[0], bool)
    freeB = np.ones(m.shape[1], bool)
    rounds, made = 0, 0
    while True:
        sub = np.where(freeA[:, None] & freeB[None, :], m, -1.0)
        rb = sub.argmax(1)
        cb = sub.argmax(0)
        acc = [(i, rb[i]) for i in range(m.shape[0]) if freeA[i] and sub[i, rb[i]] > 0 and cb[rb[i]] == i]
        rounds += 1
        if not acc:
            return rounds, made
        for i, j in acc:
            freeA[i] = False
            freeB[j] = False
        made += len(acc)


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    mt, H, W, boxes, _ = bench.WORKLOADS[wl]
    torch.cuda.set_device(0)
    body = isl_b200.Body(synth.make_flat_weights(mt, seed=0, init="torch"), mt, scale_search=bench.SCALES)
    frame = synth.synth_frame(H, W, 0)
    cand, sub = body(frame)
    torch.cuda.synchronize()
    ws = next(iter(body._work.values()))
    parts = body.njoint - 1
    counts = ws["counts"].cpu().numpy()[:parts]
    print("%s: %d candidates, %d persons; peaks per part min/mean/max %d / %.1f / %d" % (wl, len(cand), len(sub), counts.min(), counts.mean(), counts.max()))
    ps = ws["pair_score"].cpu().numpy()
    cc = ws["conn_count"].cpu().numpy()
    nl = 24 if mt == "body25" else 19
    la, lb = bench_limbs(mt)
    tot_valid = tot_pairs = tot_conn = 0
    for k in range(nl):
        nA, nB = int(counts[la[k]]), int(counts[lb[k]])
        if nA == 0 or nB == 0:
            print('  limb %2d: %3d x %3d pairs' % (k, nA, nB))
            continue
        m = ps[k, :nA * nB].reshape(nA, nB)
        valid = int((m > 0).sum())
        r, made = rounds_needed(m)
        tot_valid += valid
        tot_pairs += nA * nB
        tot_conn += int(cc[k])
        print("  limb %2d: %3d x %3d pairs, %5d with a positive score, %3d connections (device %3d), %3d rounds" % (k, nA, nB, valid, made, cc[k], r))
    print("  total: %d pairs, %d positive, %d connections" % (tot_pairs, tot_valid, tot_conn))


def bench_limbs(mt):
    if mt == "body25":
        a = [1, 1, 2, 3, 1, 5, 6, 1, 8, 9, 10, 8, 12, 13, 0, 0, 15, 16, 11, 11, 14, 14, 22, 19]
        b = [0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 24, 22, 21, 19, 23, 20]
    else:
        a = [1, 1, 2, 3, 5, 6, 1, 8, 9, 1, 11, 12, 1, 0, 14, 0, 15, 2, 5]
        b = [2, 5, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 0, 14, 16, 15, 17, 16, 17]
    return a, b


if __name__ == "__main__":
    main()
