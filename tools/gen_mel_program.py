#!/usr/bin/env python
"""Generates speechain_b200/csrc/sfe_melprog.h: the tile kernel's Mel stage as STRAIGHT-LINE code for the
filterbanks of the shipped recipes.

    python tools/gen_mel_program.py            # rewrites the header (needs torch + torchaudio, no GPU)

Why generated code: in the tile kernel a lane owns one frame and walks the bins in order with one packed FMA per bin
into an accumulator pair (even filter, odd filter); a filter is stored ("emitted") when its last bin has been
added.  WHICH bins emit is a property of the filterbank's band edges only — torchaudio's melscale_fbanks for a given
(sr, n_fft, n_mels, fmin, fmax, scale, norm), i.e. linear2mel.py:135-143 — so for the recipes' banks the emit
positions, output addresses and accumulator slots are compile-time constants: no flag tests, no branches, no address
arithmetic (a data-driven loop costs 6 instructions per bin, this one ~2).  The weights stay run-time data: the plan
uploads them in the (even, odd) layout and sfe_plan_create checks that the filterbank it was given has exactly the
band structure compiled here (otherwise the gather kernel runs).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speechain_b200.module.frontend._common import mel_filterbank  # noqa: E402

# name -> (bank id, mel_filterbank arguments)
BANKS = [
    ("16000_512_80", 1, (16000, 512, 80, 0.0, 8000.0, "slaney", True)),    # BASELINE C1/C2/C4
    ("16000_400_80", 2, (16000, 400, 80, 0.0, 8000.0, "slaney", True)),    # ASR recipes: n_fft = win_length
]


def sequential(W):
    """Same pass as melsched_detail::sequential (sfe_melsched.h): emits before each bin, filter held by the
    accumulator's low slot while the bin is added."""
    D, K = W.shape
    pre, curs = np.zeros(K, int), np.zeros(K, int)
    cur = 0
    for k in range(K):
        nz = np.nonzero(W[:, k])[0]
        if len(nz):
            assert nz.max() - nz.min() <= 1, "a bin feeds non-adjacent filters"
            need = max(0, nz.max() - (cur + 1))
            pre[k] = need
            cur += need
            assert nz.min() >= cur
        curs[k] = cur
    return pre, curs, D - cur


def build(name, W):
    D, K = W.shape
    pre, curs, tail = sequential(W)
    ng = (K + 7) // 8
    kB = 4 * ng
    assert kB < K and D % 4 == 0
    d_split = int(curs[kB])
    tail_a = int(pre[kB])
    b_off = ((kB - d_split) + 3) // 4 * 4
    c_split = d_split // 4
    rowf = 8 * ng
    assert d_split <= kB - 3 and D - d_split >= 2 and b_off + D <= rowf
    ops = []  # the program
    curA, curB = 0, d_split

    pend = {"A": [], "B": []}  # outputs of a stream waiting to leave as one 128-bit store: (filter d, slot)

    def flush_scalar(stream):
        for d, slot in pend[stream]:
            ops.append(f"E{stream}1({slot}, {d if stream == 'A' else b_off + d})")
        pend[stream].clear()

    def emit(stream, g):
        """Filter `cur` of the stream is complete.  Four consecutive outputs starting at a multiple of four leave as
        ONE 128-bit store (a row-strided 32-bit store is a 4-way bank conflict; the 128-bit one is conflict-free);
        the few outputs outside such a group are stored one by one."""
        nonlocal curA, curB
        # loads of group g + 1 are issued before group g's bins: a store must land below the first bin NOT yet loaded
        limit = (4 * (g + 2)) if stream == "A" else (kB + 4 * (g + 2))
        d = curA if stream == "A" else curB
        pos = d if stream == "A" else b_off + d
        assert pos < limit and (stream == "A" or pos >= kB)
        if d % 4 == 0:
            flush_scalar(stream)
        if d % 4 == len(pend[stream]):      # continues (or starts) an aligned group
            pend[stream].append((d, d & 1))
            ops.append(f"E{stream}({d & 1}, {d % 4})")
            if d % 4 == 3:
                ops.append(f"S{stream}({pos - 3})")
                pend[stream].clear()
        else:                               # the stream started inside a group: plain store
            ops.append(f"E{stream}1({d & 1}, {pos})")
        if stream == "A":
            curA += 1
        else:
            curB += 1

    ops.append("LD(0, 0)")
    for g in range(ng):
        if g + 1 < ng:
            ops.append(f"LD({g + 1}, {(g + 1) & 1})")
        for i in range(4):
            kA, kBk = 4 * g + i, kB + 4 * g + i
            for _ in range(int(pre[kA])):
                emit("A", g)
            assert curA == curs[kA]
            ops.append(f"FA({g & 1}, {i})")
            n = int(pre[kBk]) if (kBk < K and kBk != kB) else 0
            for _ in range(n):
                emit("B", g)
            if kBk < K:
                assert curB == curs[kBk]
            ops.append(f"FB({g & 1}, {i})")
    for _ in range(tail_a):
        emit("A", ng)
    assert curA == d_split
    flush_pending_a = list(pend["A"])
    pend["A"].clear()
    for k, (d, slot) in enumerate(flush_pending_a):  # (already taken out of the accumulator: registers -> row)
        ops.append(f"PA({k}, {d})")
    for _ in range(tail):
        emit("B", ng)
    assert curB == D and not pend["B"], "80 filters end on a group boundary"
    # A's accumulator still holds its share of filters d_split (slot d_split & 1) and d_split + 1
    ops.append(f"JOIN({d_split & 1}, {b_off + d_split})")
    for d in range(4 * c_split, d_split):
        ops.append(f"MOVE({d}, {b_off + d})")
    meta = dict(K=K, D=D, NG=ng, B_LD=kB, D_SPLIT=d_split, B_OFF=b_off, C_SPLIT=c_split, ROWF=rowf, TAIL=int(tail))
    return meta, pre, curs, ops


def main():
    out = []
    out.append("// GENERATED by tools/gen_mel_program.py — do not edit.  Straight-line Mel programs of the tile kernel\n"
               "// (sfe_tile.cuh) for the filterbanks of the shipped recipes; see the generator for the why and the how.\n"
               "// Macros: LD(g, buf) load group g's bins and weights into register set buf; FA/FB(buf, i) packed FMA of\n"
               "// bin i of the group (stream A / B); EA/EB(slot, k) move accumulator slot (0 even filter, 1 odd) to\n"
               "// output register k of the stream and clear it; SA/SB(pos) store the stream's four output registers to\n"
               "// row[pos..pos+3]; EA1/EB1(slot, pos) store the slot to row[pos] directly and clear it; PA(k, pos) store\n"
               "// output register k of stream A to row[pos]; tail (after BOTH streams): JOIN(slot, pos) add stream A's\n"
               "// leftovers to row[pos], row[pos+1]; MOVE(from, to) row[to] = row[from].\n#pragma once\n\nnamespace sfe {\nnamespace melprog {\n")
    for name, bank, args in BANKS:
        W = mel_filterbank(*args).numpy()
        meta, pre, curs, ops = build(name, W)
        out.append(f"// bank {bank}: mel_filterbank{args}\nstruct Bank{bank} {{\n")
        for k, v in meta.items():
            out.append(f"  static constexpr int {k} = {v};\n")
        out.append("};\n")
        # host-side identity check: per bin, the filter index held by the accumulator's low slot
        vals = ", ".join(str(int(c)) for c in curs)
        out.append(f"static const unsigned char kCur{bank}[{len(curs)}] = {{{vals}}};\n")
        def wrap(seq):
            lines, cur = [], "  "
            for op in seq:
                if len(cur) + len(op) + 1 > 116:
                    lines.append(cur.rstrip() + " \\\n")
                    cur = "  "
                cur += op + " "
            lines.append(cur.rstrip() + "\n")
            return "".join(lines)
        # the two streams never touch each other's bins or outputs before the tail (JOIN, MOVE): they may run on
        # two warps, the tail after both are done
        cut = next(i for i, op in enumerate(ops) if op.startswith("JOIN"))
        assert all(op.startswith(("JOIN", "MOVE")) for op in ops[cut:]) and not any(op.startswith(("JOIN", "MOVE")) for op in ops[:cut])
        out.append(f"#define SFE_MELPROG_{bank}(LD, FA, FB, EA, EB, SA, SB, EA1, EB1, PA) \\\n" + wrap(ops[:cut]) + "\n")
        out.append(f"#define SFE_MELTAIL_{bank}(JOIN, MOVE) \\\n" + wrap(ops[cut:]) + "\n")
    out.append("}  // namespace melprog\n}  // namespace sfe\n")
    path = os.path.join(ROOT, "speechain_b200", "csrc", "sfe_melprog.h")
    with open(path, "w") as f:
        f.write("".join(out))
    print("wrote", path)


if __name__ == "__main__":
    main()
