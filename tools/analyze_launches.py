"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: totals per kernel, per network plan and,
for the largest body / hand plan, per layer with its algorithmic TFLOP/s."""
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isl_b200  # noqa: E402,F401
from isl_b200 import nets  # noqa: E402


def ns(row):
    return float(row['Metric Value'].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}[row['Metric Unit']]


def main(path, body_kind, body_hw, hand_hw, n_body, n_hand):
    rows = list(csv.DictReader([l for l in open(path) if not l.startswith('==')]))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r'\(.*', '', r['Kernel Name']).replace('islpose::', '').replace('<unnamed>::', '')
        tot[name][0] += 1
        tot[name][1] += ns(r)
    allt = sum(v[1] for v in tot.values())
    print("== per kernel (all %d launches, %.1f ms)" % (len(rows), allt / 1e6))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:12]:
        print("  %-52s n=%5d %9.3f ms %5.1f%%  avg %8.1f us" % (k[:52], v[0], v[1] / 1e6, 100 * v[1] / allt, v[1] / v[0] / 1e3))
    bystream = collections.defaultdict(list)
    for r in rows:
        if any(t in r['Kernel Name'] for t in ('conv_umma', 'conv_first')):
            bystream[r['Stream']].append(r)
    last = {}
    for st, lst in bystream.items():
        cur = None
        for r in lst:
            if 'conv_first' in r['Kernel Name']:
                cur = []
                last[(st, r['Grid Size'])] = cur
            elif cur is not None:
                cur.append(r)
    print("== last replay of every plan")
    plans = []
    for (st, g), p in last.items():
        t = sum(ns(r) for r in p)
        plans.append((t, st, len(p), p))
        print("  stream %s: %3d launches %8.3f ms" % (st, len(p), t / 1e6))
    for kind, hw, n, length in ((body_kind, body_hw, n_body, None), ("hand", hand_hw, n_hand, 56)):
        prog = nets.build_program(kind)
        steps = [s for s in prog.steps if s[0] == 'conv']   # the pools are fused into their producers
        cands = [p for p in plans if p[2] == len(steps)]
        if not cands:
            continue
        t, st, _, plan = max(cands)
        print("== %s plan n=%d %dx%d: %.3f ms" % (kind, n, hw[0], hw[1], t / 1e6))
        groups = collections.OrderedDict()
        tf = 0
        for s, r in zip(steps, plan):
            if s[0] == 'pool':
                key, f = 'pool', 0
            else:
                d = s[1]
                level = prog.bufs[d['src'][0]][1]
                cin = 3 if d['first'] else (sum(1 for c in d['chan_map'] if c is not None) if d['chan_map'] else d['src'][2])
                f = 2 * cin * d['cout'] * d['k'] ** 2 * (hw[0] >> level) * (hw[1] >> level) * n
                key = re.sub(r'_L[12]$', '', re.sub(r'stage\d', 'stageX', d['layer']))
                key = re.sub(r'Mconv\d_stageX_L\d_(\d)', r'dense_\1', key)
            g = groups.setdefault(key, [0, 0.0, 0.0, set()])
            g[0] += 1
            g[1] += f
            g[2] += ns(r)
            g[3].add(re.sub(r'.*::', '', re.sub(r'\(.*', '', r['Kernel Name']))[10:22] + r['Grid Size'])
            tf += f
        for k, (c, f, tt, kinds) in groups.items():
            print("  %-20s x%2d %9.1f us %5.1f%% %7.1f TFLOP/s  %s" % (k, c, tt / 1e3, 100 * tt / t, f / tt / 1e3 if tt else 0,
                                                                     sorted(kinds)[0][:40]))
        print("  total %.1f GFLOP -> %.1f TFLOP/s" % (tf / 1e9, tf / t / 1e3))


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("csv", help="ncu --metrics gpu__time_duration.sum --csv launch list")
    ap.add_argument("--body", default="coco", choices=["coco", "body25"])
    ap.add_argument("--body-hw", type=int, nargs=2, default=[736, 984], help="largest body network input (h w)")
    ap.add_argument("--hand-hw", type=int, default=736, help="largest hand network input (square)")
    ap.add_argument("--n-body", type=int, default=8, help="frames per body replay")
    ap.add_argument("--n-hand", type=int, default=16, help="crops per hand replay")
    a = ap.parse_args()
    main(a.csv, a.body, tuple(a.body_hw), (a.hand_hw, a.hand_hw), a.n_body, a.n_hand)
