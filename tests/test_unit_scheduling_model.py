"""CPU model of the fused kernel's unit enumeration (speechain_b200/csrc/sfe_fused.cuh, producer warp).

The producer lanes of a CTA take chunks of the CTA's range of the *cost axis* from a shared counter and turn a
chunk start into (utterance, unit) by the arithmetic restated here line by line (cost table, forward seek,
`settle`, runs of pad units).  The model deals the chunks to the lanes in random order and checks the property
the kernel relies on: whatever the interleaving, every unit of the batch is handed out exactly once (real units
one per descriptor, pad units in runs), for ragged and empty utterances, odd frame counts, batches beyond the
cost table and grids larger than the work.  It guards the integer arithmetic; the GPU tests guard the kernel.
"""
import random

FPW, RC = 2, 8
def run(lens_frames, F, compact, tabled, grid, nw, rng):
    B = len(lens_frames); upu = (F + FPW - 1) // FPW
    flen = [max(0, min(F, f)) for f in lens_frames]
    cost = [RC * ((fl + FPW - 1) // FPW) + (0 if compact else upu - (fl + FPW - 1) // FPW) for fl in flen]
    s_cum = [0]
    for c in cost: s_cum.append(s_cum[-1] + c)
    ctot = s_cum[B] if tabled else B * upu
    real_step = RC if tabled else 1
    seen = {}
    for cta in range(grid):
        cta_lo = ctot * cta // grid; cta_hi = ctot * (cta + 1) // grid
        s_next = [0]
        class Lane: pass
        lanes = []
        for i in range(nw):
            L = Lane(); L.b = -1; L.tu = 0; L.flen = 0; L.nunit = 0; L.c = 0; L.c_hi = 0; L.first = True; L.live = True
            lanes.append(L)
        def load(L):
            L.flen = flen[L.b]; L.nunit = ((L.flen + FPW - 1) // FPW) if (compact and tabled) else upu
        def settle(L):
            while L.tu >= L.nunit and L.b + 1 < B:
                L.b += 1; L.tu = 0; load(L)
                if tabled: L.c = s_cum[L.b]
        def serve(L):
            have = False
            if (not L.first) and L.tu * FPW >= L.flen:
                L.c += 1; L.tu += 1; settle(L)
                have = L.tu < L.nunit and L.c < L.c_hi
            L.first = False
            while not have:
                k = s_next[0]; s_next[0] += 1
                lo64 = cta_lo + k * real_step
                if lo64 >= cta_hi: L.live = False; break
                c0 = lo64; L.c_hi = min(c0 + real_step, cta_hi)
                if tabled:
                    lo = L.b
                    if lo < 0 or s_cum[lo] > c0: lo = 0
                    if lo + 1 < B and s_cum[lo + 1] <= c0:
                        hi = B - 1; lo += 1
                        while lo < hi:
                            mid = (lo + hi + 1) >> 1
                            if s_cum[mid] <= c0: lo = mid
                            else: hi = mid - 1
                    if lo != L.b: L.b = lo; load(L)
                    nreal = (L.flen + FPW - 1) // FPW
                    off = c0 - s_cum[L.b]; real_cost = RC * nreal
                    if off <= real_cost:
                        L.tu = int((off + RC - 1) / RC) if off + RC - 1 < 0 else (off + RC - 1) // RC   # C++ truncation vs floor only differs for negatives
                        L.c = c0 - off + L.tu * RC
                    else:
                        L.tu = nreal + (off - real_cost); L.c = c0
                else:
                    nb = c0 // upu
                    if nb >= B: nb = B - 1
                    L.tu = c0 - nb * upu; L.c = c0
                    if nb != L.b: L.b = nb; load(L)
                settle(L)
                have = L.tu < L.nunit and L.c < L.c_hi
            if not L.live: return
            if L.tu * FPW < L.flen:
                key = (L.b, L.tu); assert key not in seen, ("dup real", key); seen[key] = "real"
            else:
                n = min(L.c_hi - L.c, L.nunit - L.tu); assert n >= 1
                for j in range(n):
                    key = (L.b, L.tu + j); assert key not in seen, ("dup pad", key); seen[key] = "pad"
                L.tu += n - 1; L.c += n - 1
        alive = lanes[:]
        steps = 0
        while alive:
            L = rng.choice(alive); serve(L)
            if not L.live: alive.remove(L)
            steps += 1; assert steps < 10**7
    # expected units
    exp = {}
    for b in range(B):
        nreal = (flen[b] + FPW - 1) // FPW
        n = nreal if (compact and tabled) else upu
        for tu in range(n): exp[(b, tu)] = "real" if tu * FPW < flen[b] else "pad"
    if compact and not tabled:
        # pad units may be enumerated (and skipped); real ones must match
        assert {k for k, v in exp.items() if v == "real"} == {k for k, v in seen.items() if v == "real"}
        assert set(seen) <= set(exp)
    assert seen == exp or (compact and not tabled), (len(seen), len(exp), set(exp) ^ set(seen))
def test_every_unit_is_handed_out_exactly_once():
  rng = random.Random(0)
  for trial in range(400):
      B = rng.choice([1, 2, 3, 5, 17, 64, 300])
      F = rng.choice([1, 2, 3, 7, 20, 101])
      lens = [rng.choice([0, 0, 1, 2, F // 2, F - 1, F, F + 3, rng.randint(0, F + 2)]) for _ in range(B)]
      if rng.random() < 0.3: lens = sorted(lens, reverse=True)
      compact = rng.random() < 0.5
      tabled = B <= 256
      grid = rng.choice([1, 2, 7, 148]); nw = rng.choice([1, 3, 7, 23])
      run(lens, F, compact, tabled, grid, nw, rng)
