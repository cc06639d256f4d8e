"""Model of the goods market's commit rounds (development aid): counts rounds per period for commit rules,
on real pre-market states taken from the oracle.  Every rule must reproduce the sequential outcome.

rules: R0 = commit the prefix up to the first lane that exhausts its seller (the kernel until round 2);
       R1 = also commit lanes above it whose terminal cannot be reached by an uncommitted lower lane
            (terminal not in the remaining candidates of any uncertain lower lane; cascade to fixpoint).
window = lanes per chunk (32 = one warp; 64..256 = several warps committing one super-chunk together)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from oracle import oracle as orc


def premarket(W, F, N, seed, t):
    e = orc.OracleEconomy(W, F, N, seed=seed, economy=0)
    for _ in range(t):
        e.step()
    rec = e.fill_tape()
    off = orc.tape_offsets(W, F, N, 5, 2, 5)
    H = W + F + N
    order = rec[off["goods_order"]: off["goods_order"] + 4 * H].view(np.int32).copy()
    cand = rec[off["goods_cand"]: off["goods_cand"] + 4 * H * 5].view(np.int32).reshape(H, 5).copy()
    e.set_stop_phase(19); e.step()
    return order, cand, e.f64("budget").copy(), e.f64("c_Ystock").copy(), e.f64("c_P").copy()


def sequential(order, cand, budget, stock, P):
    stock = stock.copy(); left = np.zeros(len(budget))
    for h in order:
        b = budget[h]
        if not b > 0: continue
        for c in sorted(cand[h], key=lambda c: (P[c], list(cand[h]).index(c))):
            if not b > 0: break
            ys = stock[c]
            if ys > 0:
                d = b * (1.0 / P[c])
                if ys > d: stock[c] = ys - d; b = 0.0
                else: b = b - ys * P[c]; stock[c] = 0.0
        left[h] = b
    return stock, left


def rounds(order, cand, budget, stock, P, rule, window):
    stock = stock.copy(); left = np.zeros(len(budget))
    H = len(order); n_rounds = 0; n_chunks = 0; committed_hist = []
    for c0 in range(0, H, window):
        hs = order[c0:c0 + window]; n = len(hs)
        b = [budget[h] for h in hs]
        lists = [sorted(cand[h], key=lambda c, h=h: (P[c], list(cand[h]).index(c))) for h in hs]
        pos = [0] * n                      # next candidate to pop
        term = [None if b[i] > 0 else -1 for i in range(n)]   # None = must walk, -1 = done, >= 0 terminal
        pend = [b[i] > 0 for i in range(n)]
        n_chunks += 1
        while any(pend):
            n_rounds += 1
            for i in range(n):             # walk
                while pend[i] and term[i] is None:
                    if pos[i] >= len(lists[i]): term[i] = -1; break
                    s = lists[i][pos[i]]; pos[i] += 1
                    if stock[s] > 0: term[i] = s
            # replay in lane order per seller
            ys_at = {}; status = [None] * n; ys_seen = [0.0] * n
            for i in range(n):
                if not pend[i]: continue
                if term[i] == -1: status[i] = "none"; continue
                s = term[i]
                ys = ys_at.get(s, stock[s])
                if ys is None: status[i] = "sold"; continue
                d = b[i] * (1.0 / P[s])
                ys_seen[i] = ys
                if ys > d: status[i] = "full"; ys_at[s] = ys - d
                else: status[i] = "exh"; ys_at[s] = None
            movers = [i for i in range(n) if status[i] in ("exh", "sold")]
            first = movers[0] if movers else n
            if rule == "R0":
                commit = [i for i in range(n) if pend[i] and i <= first and status[i] != "sold"]
            else:
                # uncertain lanes: movers, then cascade on remaining-candidate sets
                unc = set(movers); changed = True
                while changed:
                    changed = False
                    danger = {}
                    for i in sorted(unc):
                        for s in lists[i][pos[i]:]:
                            danger.setdefault(s, i)
                            danger[s] = min(danger[s], i)
                    for j in range(n):
                        if pend[j] and j not in unc and status[j] == "full" and danger.get(term[j], n) < j:
                            unc.add(j); changed = True
                if rule == "R3":
                    # no mover analysis at all: a lane commits unless its terminal is on the remaining list of ANY lower pending lane
                    danger = {}
                    for i in range(n):
                        if pend[i] and term[i] is not None and term[i] >= 0:
                            for s_ in lists[i][pos[i]:]:
                                danger[s_] = min(danger.get(s_, n), i)
                    commit = [i for i in range(n) if pend[i] and (status[i] == "none" or (status[i] in ("full", "exh") and not danger.get(term[i], n) < i))]
                if rule == "R4":
                    unc = set(movers)
                    danger = {}
                    for i in sorted(unc):
                        for s_ in lists[i][pos[i]:]:
                            danger[s_] = min(danger.get(s_, n), i)
                    newly = [j for j in range(n) if pend[j] and j not in unc and status[j] == "full" and danger.get(term[j], n) < j]
                    limit = min(newly) if newly else n
                    commit = [i for i in range(n) if pend[i] and i < limit and ((status[i] in ("full", "none") and i not in unc)
                              or (status[i] == "exh" and not danger.get(term[i], n) < i))]
                    commit += [i for i in range(n) if pend[i] and i >= limit and status[i] == "none"]
                # a mover is itself certain unless its terminal is endangered by a LOWER uncertain lane
                def endangered(j):
                    return any(i < j and term[j] in lists[i][pos[i]:] for i in unc if i != j)
                if rule not in ("R4", "R3"):
                    commit = [i for i in range(n) if pend[i] and status[i] in ("full", "none") and i not in unc]
                if rule in ("R4", "R3"):
                    pass
                elif rule == "R1":
                    if first < n and status[first] == "exh": commit.append(first)
                else:
                    commit += [i for i in movers if status[i] == "exh" and not endangered(i)]
            committed_hist.append(len(commit))
            for i in commit:
                if status[i] == "none": pend[i] = False
                elif status[i] == "full":
                    s = term[i]; stock[s] = ys_seen[i] - b[i] * (1.0 / P[s]); b[i] = 0.0; pend[i] = False
                else:   # a mover that exhausts its seller
                    s = term[i]; b[i] = b[i] - ys_seen[i] * P[s]; stock[s] = 0.0
                    term[i] = None
                    if not b[i] > 0: pend[i] = False; term[i] = -1
            # lanes whose terminal is now empty walk on
            for i in range(n):
                if pend[i] and term[i] is not None and term[i] >= 0 and not stock[term[i]] > 0: term[i] = None
        for i in range(n): left[hs[i]] = b[i]
    return stock, left, n_rounds, n_chunks, committed_hist


WINDOWS = tuple(int(x) for x in sys.argv[1:]) or (32, 128, 256)


if __name__ == "__main__":
    for (W, F, N, t) in ((1000, 100, 20, 320), (1000, 100, 20, 350), (10000, 1000, 200, 120)):
        order, cand, budget, stock, P = premarket(W, F, N, 1, t)
        s_ref, l_ref = sequential(order, cand, budget, stock, P)
        print(f"W={W} F={F} t={t}: buyers={int((budget > 0).sum())} sold-out sellers={int((s_ref == 0).sum())}/{F}")
        for window in WINDOWS:
            for rule in ("R2", "R3"):
                s, l, nr, nc, hist = rounds(order, cand, budget, stock, P, rule, window)
                ok = np.array_equal(s, s_ref) and np.array_equal(l, l_ref)
                print(f"  window {window:3d} {rule}: rounds {nr:5d} chunks {nc:4d} rounds/chunk {nr / nc:.2f} exact={ok}")
