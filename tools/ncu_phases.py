#!/usr/bin/env python
"""Time share / instruction count per phase of the sweep from an ncu capture
(uses tools/ncu_lines.py and marker comments in riccati_group.cuh)."""
import collections
import glob
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as nl

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, obj, krx, knots = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    fn, lm = nl.sass_linemap(obj, krx)
    data = nl.ncu_source(rep)
    src = open(os.path.join(ROOT, "aligator_b200/csrc/riccati_group.cuh")).read().splitlines()

    def find(pat):
        for i, l in enumerate(src, 1):
            if pat in l:
                return i
    marks = [(1, "helpers(dot/ld)"), (find("AB2_D bool bk_factor_group"), "bk_smem"),
             (find("AB2_D void bk_solve_column"), "solve(general)"),
             (find("template <int N> struct FastFactor"), "fastBK+solve"),
             (find("template <int N> struct RegFactor"), "regBK"),
             (find("AB2_D void bk_solve_vec_group"), "solve_vec"),
             (find("AB2_D void stage_loop_mma"), "mma:setup+load"),
             (find("// (1) W = V' M   (+ vx'"), "mma:(1)W"),
             (find("// (2) H = H0 + M^T W"), "mma:(2)H"),
             (find("// (3) control rows of H: [Shat^T | rhat] -> X"), "mma:(3)dump+BK+solve"),
             (find("// fragments of KK = [K k] (rows >= NK are zero)"), "mma:(4)Ahat"),
             (find("// (5) [Vxx vx] = [Qhat qhat] + Shat KK"), "mma:(5)Vxx+stores"),
             (find("AB2_D void riccati_group_sweep"), "setup"),
             (find("terminal knot (nu = 0)"), "terminal"),
             (find("stage knots N-1 .. 0"), "stage:load+A"), (find("(B) H[:,j] = H0"), "stage:B"),
             (find("(C) Bunch-Kaufman of the reduced"), "stage:C/D glue"),
             (find("(E) cost-to-go"), "stage:E+stores"), (find("initial stage: proximal"), "kkt0"),
             (find("forward rollout: riccati-kernel"), "forward:x"),
             (find("Pass 2 -- the parallel part"), "forward:lambda")]
    marks = sorted(m for m in marks if m[0])

    def region(line):
        r = "?"
        for a, n in marks:
            if line >= a:
                r = n
        return r
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = 0
    for off, sass, smp, nexe, st in data:
        ent = lm.get(off)
        if ent and ent[0] and ent[0][0] == "riccati_group.cuh":
            key = region(ent[0][1])
        elif ent and ent[0]:
            key = "launch.cuh(ctx)"
        else:
            key = "?"
        agg[key][0] += smp
        agg[key][1] += nexe
        agg[key][2].update(st)
        tot += smp
    ti = sum(v[1] for v in agg.values())
    print("kernel", fn[:100])
    print("total samples %d, warp-instr/knot %.1f" % (tot, ti / knots))
    for k, (s, n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print("%-18s %6.2f%%  instr/knot %7.1f   %s" % (
            k, 100.0 * s / tot, n / knots,
            ", ".join("%s %.0f%%" % (a, 100.0 * b / max(s, 1)) for a, b in st.most_common(4))))


if __name__ == "__main__":
    main()
