"""Developer model of the VW_WAVEFRONT synchronisation (k_fused_analysis): random interleavings of the warps, every
shared-memory cell tagged with the level that wrote it, every read checked against the level it must see.  Confirms that
waiting for warps w+1 and w+2 is sufficient (RAW on the previous level's outputs and WAR on the buffer being recycled) and
that dropping either wait is caught.  Pure Python, no GPU."""
import random
import sys


def run(L, d0, nlev, T, nthreads=256, R=9, waits=(1, 2), seed=0, trials=200):
    hexact = (L - 1) * d0 * ((1 << nlev) - 1)
    HT = (hexact + 1) & ~1
    PP = HT + T
    nw = nthreads // 32
    rng = random.Random(seed)
    for _ in range(trials):
        # tag[buf][i] = level index whose outputs the cell holds (-1 = the input tile, valid everywhere)
        tag = [[-1] * PP, [None] * PP]
        prog = [0] * nw
        lev_of = [0] * nw                      # next level each warp will run
        lo = [HT - hexact]
        for lev in range(nlev):
            lo.append(lo[-1] + (L - 1) * (d0 << lev))
        while any(l < nlev for l in lev_of):
            ready = [w for w in range(nw) if lev_of[w] < nlev and
                     all(w + k >= nw or prog[w + k] >= lev_of[w] for k in waits)]
            if not ready:
                return "deadlock"
            w = rng.choice(ready)
            lev = lev_of[w]
            d = d0 << lev
            last = lev + 1 == nlev
            ra = HT if last else lo[lev + 1]
            cur, nxt = tag[lev & 1], tag[(lev + 1) & 1]
            outs = []
            for lane in range(32):
                tid = w * 32 + lane
                c, ph = tid // d, tid % d
                base = ra + c * R * d + ph
                for r in range(R):
                    p = base + r * d
                    if p >= PP:
                        break
                    for k in range(L):                 # reads of the previous level's outputs
                        if cur[p - k * d] != lev - 1:
                            return f"RAW/WAR violation: warp {w} level {lev} read cell {p - k * d} holding {cur[p - k * d]}"
                    outs.append(p)
            for p in outs:
                nxt[p] = lev
            prog[w] = lev + 1
            lev_of[w] = lev + 1
    return "ok"


if __name__ == "__main__":
    bad = 0
    for (L, d0, nlev, T) in ((8, 1, 4, 2048), (2, 1, 4, 2048), (8, 4, 3, 2048), (4, 2, 4, 2200), (8, 8, 3, 1900), (6, 1, 5, 2000)):
        full = run(L, d0, nlev, T)
        no1 = run(L, d0, nlev, T, waits=(2,))
        no2 = run(L, d0, nlev, T, waits=(1,))
        print(f"L={L} d0={d0} nlev={nlev} T={T}: waits(w+1,w+2) -> {full};  without w+1 -> {no1[:40]};  without w+2 -> {no2[:40]}")
        bad += full != "ok"
    sys.exit(1 if bad else 0)
