"""Offline model of the shared-memory gather of kLJPruned: builds per-lane partner lists for a C2-like liquid the way
pruned.cu does (towers, z-sorted 32-slot chunks, 2x2x2 tiles, compact staged indices) and counts LDS wavefronts per
row for different orderings of each lane's list. LDS.128 (x, y): a quarter-warp per wavefront, bank group = c mod 8;
LDS.64 (z): a half-warp per wavefront, bank pair = c mod 16; equal addresses broadcast."""
import sys
import numpy as np
from scipy.spatial import cKDTree

rng = np.random.default_rng(0)
npd = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rho, rc, skin, M = 0.8442, 2.5, 0.3, 32
sp = rho ** (-1 / 3)
L = npd * sp
g = (np.arange(npd) + 0.5) * sp
pos = np.stack([a.ravel() for a in np.meshgrid(g, g, g, indexing="ij")], axis=1) + rng.uniform(-0.15, 0.15, (npd ** 3, 3))
pos %= L
n = len(pos)
side = (M / rho) ** (1 / 3)
nt = int(np.ceil(L / side))
side = L / nt
tx = np.minimum((pos[:, 0] / side).astype(int), nt - 1)
ty = np.minimum((pos[:, 1] / side).astype(int), nt - 1)
tower = tx + ty * nt
order = np.lexsort((pos[:, 2], tower))
pos, tower, tx, ty = pos[order], tower[order], tx[order], ty[order]
# slot within tower, chunk index
start = np.searchsorted(tower, np.arange(nt * nt))
rank = np.arange(n) - start[tower]
chunk = rank // 32
tree = cKDTree(pos, boxsize=L)
pairs = tree.query_pairs(rc + skin, output_type="ndarray")
nbr = [[] for _ in range(n)]
for a, b in pairs:
    nbr[a].append(b)
    nbr[b].append(a)
# tiles: 2x2 towers x 2 chunks
tile_of = (tx // 2) + (ty // 2) * ((nt + 1) // 2) + (chunk // 2) * (((nt + 1) // 2) ** 2)
warp_of = tile_of * 8 + ((tx & 1) + 2 * (ty & 1)) * 2 + (chunk & 1)
lane_of = rank % 32


def wavefronts(rows):
    """rows: (T, 32) compact indices, -1 = padding (conflict-free sentinel). returns (xy wavefronts, z wavefronts)"""
    wxy = wz = 0
    for r in rows:
        for q in range(4):
            c = r[8 * q:8 * q + 8]
            c = np.unique(c[c >= 0])
            wxy += max(1, np.bincount(c % 8, minlength=8).max()) if len(c) else 1
        for hlf in range(2):
            c = r[16 * hlf:16 * hlf + 16]
            c = np.unique(c[c >= 0])
            wz += max(1, np.bincount(c % 16, minlength=16).max()) if len(c) else 1
    return wxy, wz


def order_current(lst, lane):
    return lst


def order_rot16(lst, lane):
    lst = np.asarray(lst)
    return lst[np.argsort((lst - lane) % 16, kind="stable")]


def order_rot8(lst, lane):
    lst = np.asarray(lst)
    return lst[np.argsort((lst - lane) % 8, kind="stable")]


def greedy_group(lists, T, G, pick="most"):
    """G lanes, classes mod G: per slot, lanes in descending remaining length pick a free class; leftovers go into the
    lane's free slots afterwards (conflicts allowed there). Returns rows (T, G)."""
    rem = [dict() for _ in range(G)]
    for l, lst in enumerate(lists):
        for c in lst:
            rem[l].setdefault(c % G, []).append(c)
    out = -np.ones((T, G), dtype=int)
    left = [len(x) for x in lists]
    for t in range(T):
        taken = set()
        for l in sorted(range(G), key=lambda q: -left[q]):
            best = None
            if pick == "most":
                for cls, q in rem[l].items():
                    if q and cls not in taken and (best is None or len(q) > len(rem[l][best])):
                        best = cls
            else:  # rotating first-fit starting at (lane + t) mod G
                for k in range(G):
                    cls = (l + t + k) % G
                    if cls not in taken and rem[l].get(cls):
                        best = cls
                        break
            if best is not None:
                out[t, l] = rem[l][best].pop()
                taken.add(best)
                left[l] -= 1
    nover = 0
    for l in range(G):
        rest = [c for q in rem[l].values() for c in q]
        nover += len(rest)
        free = [t for t in range(T) if out[t, l] < 0]
        for t, c in zip(free, rest):
            out[t, l] = c
    return out, nover


def proposal_group(lists, T, G, rounds):
    """Parallel variant: per slot every lane proposes its first available class at a rotating start; among lanes with
    equal proposals the one with most entries left wins; losers retry `rounds - 1` times among classes not yet taken."""
    rem = [dict() for _ in range(G)]
    for l, lst in enumerate(lists):
        for c in lst:
            rem[l].setdefault(c % G, []).append(c)
    out = -np.ones((T, G), dtype=int)
    left = [len(x) for x in lists]
    for t in range(T):
        taken = set()
        pending = [l for l in range(G) if left[l] > 0]
        for rd in range(rounds):
            props = {}
            for l in pending:
                for k in range(G):
                    cls = (l + t + k) % G
                    if cls not in taken and rem[l].get(cls):
                        props.setdefault(cls, []).append(l)
                        break
            nxt = []
            for cls, ls in props.items():
                w = max(ls, key=lambda q: (left[q], -q))
                out[t, w] = rem[w][cls].pop()
                left[w] -= 1
                nxt += [q for q in ls if q != w]
            taken |= set(props.keys())
            pending = nxt
            if not pending:
                break
    nover = 0
    for l in range(G):
        rest = [c for q in rem[l].values() for c in q]
        nover += len(rest)
        free = [t for t in range(T) if out[t, l] < 0]
        for t, c in zip(free, rest):
            out[t, l] = c
    return out, nover


def deadline_group(lists, T, G, rounds):
    """As proposal_group, but single pass: a lane that lost every round idles only while it has slack (entries left <
    slots left); without slack it places any remaining entry at once and accepts the conflict."""
    rem = [dict() for _ in range(G)]
    for l, lst in enumerate(lists):
        for c in lst:
            rem[l].setdefault(c % G, []).append(c)
    out = -np.ones((T, G), dtype=int)
    left = [len(x) for x in lists]
    for t in range(T):
        taken = set()
        pending = [l for l in range(G) if left[l] > 0]
        for rd in range(rounds):
            props = {}
            for l in pending:
                for k in range(G):
                    cls = (l + t + k) % G
                    if cls not in taken and rem[l].get(cls):
                        props.setdefault(cls, []).append(l)
                        break
                else:
                    props.setdefault(-1 - l, []).append(l)  # nothing available
            nxt = []
            for cls, ls in props.items():
                if cls < 0:
                    nxt += ls
                    continue
                w = max(ls, key=lambda q: (left[q], -q))
                out[t, w] = rem[w][cls].pop()
                left[w] -= 1
                nxt += [q for q in ls if q != w]
            taken |= set(c for c in props.keys() if c >= 0)
            pending = nxt
            if not pending:
                break
        for l in pending:
            if left[l] >= T - t:  # no slack: place anything
                for k in range(G):
                    cls = (l + t + k) % G
                    if rem[l].get(cls):
                        out[t, l] = rem[l][cls].pop()
                        left[l] -= 1
                        break
    assert sum(left) == 0, left
    return out, 0


def latin_rows(lists, T, G=8):
    """Lane-local: lane l visits class (l + t) mod G at slot t; the j-th entry of class k sits at slot ((k - l) mod G) + G j
    while that is < T; surplus entries go to the lane's empty slots in order (may conflict)."""
    out = -np.ones((T, len(lists)), dtype=int)
    for l, lst in enumerate(lists):
        cnt = [0] * G
        surplus = []
        for c in lst:
            k = c % G
            t = ((k - l) % G) + G * cnt[k]
            if t < T:
                out[t, l] = c
                cnt[k] += 1
            else:
                surplus.append(c)
        free = [t for t in range(T) if out[t, l] < 0]
        assert len(free) >= len(surplus)
        for t, c in zip(free, surplus):
            out[t, l] = c
    return out


def latin_open(lists, T, G=16):
    """Latin layout without a surplus buffer: an entry whose preferred slot is taken or beyond T goes to the lane's last
    free slot (open addressing from the end); entries are placed in list order."""
    out = -np.ones((T, len(lists)), dtype=int)
    for l, lst in enumerate(lists):
        cnt = [0] * G
        spill = T - 1
        for c in lst:
            k = c % G
            t = ((k - l) % G) + G * cnt[k]
            cnt[k] += 1
            if t >= T or out[t, l] >= 0:
                while out[spill, l] >= 0:
                    spill -= 1
                t = spill
            out[t, l] = c
    return out


def greedy_half(lists16, T):
    """16 lanes: per slot, lanes in descending remaining length pick a free class (mod 16) with most remaining entries"""
    rem = [dict() for _ in range(16)]
    for l, lst in enumerate(lists16):
        for c in lst:
            rem[l].setdefault(c % 16, []).append(c)
    out = -np.ones((T, 16), dtype=int)
    left = [len(x) for x in lists16]
    for t in range(T):
        taken = set()
        for l in sorted(range(16), key=lambda q: -left[q]):
            best = None
            for cls, q in rem[l].items():
                if q and cls not in taken and (best is None or len(q) > len(rem[l][best])):
                    best = cls
            if best is not None:
                out[t, l] = rem[l][best].pop()
                taken.add(best)
                left[l] -= 1
    overflow = sum(left)
    return out, overflow


tiles = np.unique(tile_of)
sel = rng.choice(tiles, size=min(40, len(tiles)), replace=False)
res = {k: [0, 0, 0] for k in ("current", "rot16", "rot8", "greedy16", "g16fill", "g16fill_rot", "g8fill", "g8fill_rot", "p8r2", "p8r3", "p8r4", "p8r2_singlez", "d8r1", "d8r2", "d8r2_singlez", "d8r1_singlez", "latin8", "latin8_singlez", "latin16", "latin16_open")}
leftover = {k: 0 for k in res}


def wavefronts_soa3(rows):
    """x, y, z as three LDS.64 from separate arrays: each a half-warp per wavefront, bank pair = c mod 16"""
    w = 0
    for r in rows:
        for hlf in range(2):
            c = r[16 * hlf:16 * hlf + 16]
            c = np.unique(c[c >= 0])
            w += max(1, np.bincount(c % 16, minlength=16).max()) if len(c) else 1
    return 2 * w, w


def wavefronts_q8(rows):
    """(x, y) LDS.128 and z LDS.64 from a doubled, quarter-interleaved z array: both depend on c mod 8 per quarter"""
    wxy = 0
    for r in rows:
        for q in range(4):
            c = r[8 * q:8 * q + 8]
            c = np.unique(c[c >= 0])
            wxy += max(1, np.bincount(c % 8, minlength=8).max()) if len(c) else 1
    return wxy, wxy / 2
tot_overflow = 0
for t in sel:
    members = np.where(tile_of == t)[0]
    staged = sorted(set(j for i in members for j in nbr[i]))
    cidx = {j: k for k, j in enumerate(staged)}
    for w in np.unique(warp_of[members]):
        wm = members[warp_of[members] == w]
        lists = [[] for _ in range(32)]
        for i in wm:
            lists[lane_of[i]] = [cidx[j] for j in sorted(nbr[i])]
        T = (max(len(x) for x in lists) + 3) // 4 * 4
        for name, fn in (("current", order_current), ("rot16", order_rot16), ("rot8", order_rot8)):
            rows = -np.ones((T, 32), dtype=int)
            for l in range(32):
                o = fn(lists[l], l)
                rows[:len(o), l] = o
            a, b = wavefronts(rows)
            res[name][0] += a
            res[name][1] += b
            res[name][2] += T
        rows = -np.ones((T, 32), dtype=int)
        for hlf in range(2):
            o, ov = greedy_half(lists[16 * hlf:16 * hlf + 16], T)
            rows[:, 16 * hlf:16 * hlf + 16] = o
            tot_overflow += ov
        a, b = wavefronts(rows)
        res["greedy16"][0] += a
        res["greedy16"][1] += b
        res["greedy16"][2] += T
        for name, G, pick, wf in (("g16fill", 16, "most", wavefronts_soa3), ("g16fill_rot", 16, "rot", wavefronts_soa3),
                                  ("g8fill", 8, "most", wavefronts_q8), ("g8fill_rot", 8, "rot", wavefronts_q8)):
            rows = -np.ones((T, 32), dtype=int)
            for g0 in range(0, 32, G):
                o, ov = greedy_group(lists[g0:g0 + G], T, G, pick)
                rows[:, g0:g0 + G] = o
                leftover[name] += ov
            a, b = wf(rows)
            res[name][0] += a
            res[name][1] += b
            res[name][2] += T
        for name, rounds in (("p8r2", 2), ("p8r3", 3), ("p8r4", 4)):
            rows = -np.ones((T, 32), dtype=int)
            for g0 in range(0, 32, 8):
                o, ov = proposal_group(lists[g0:g0 + 8], T, 8, rounds)
                rows[:, g0:g0 + 8] = o
                leftover[name] += ov
            a, b = wavefronts_q8(rows)
            res[name][0] += a
            res[name][1] += b
            res[name][2] += T
            if name == "p8r2":
                a, b = wavefronts(rows)
                res["p8r2_singlez"][0] += a
                res["p8r2_singlez"][1] += b
                res["p8r2_singlez"][2] += T
        for name, rounds in (("d8r1", 1), ("d8r2", 2)):
            rows = -np.ones((T, 32), dtype=int)
            for g0 in range(0, 32, 8):
                o, ov = deadline_group(lists[g0:g0 + 8], T, 8, rounds)
                rows[:, g0:g0 + 8] = o
            a, b = wavefronts_q8(rows)
            res[name][0] += a
            res[name][1] += b
            res[name][2] += T
            a, b = wavefronts(rows)
            res[name + "_singlez"][0] += a
            res[name + "_singlez"][1] += b
            res[name + "_singlez"][2] += T
        rows = latin_rows(lists, T)
        a, b = wavefronts_q8(rows)
        res["latin8"][0] += a; res["latin8"][1] += b; res["latin8"][2] += T
        a, b = wavefronts(rows)
        res["latin8_singlez"][0] += a; res["latin8_singlez"][1] += b; res["latin8_singlez"][2] += T
        rows = latin_rows(lists, T, 16)
        a, b = wavefronts(rows)
        res["latin16"][0] += a; res["latin16"][1] += b; res["latin16"][2] += T
        rows = latin_open(lists, T, 16)
        a, b = wavefronts(rows)
        res["latin16_open"][0] += a; res["latin16_open"][1] += b; res["latin16_open"][2] += T
for k, (a, b, T) in res.items():
    print(f"{k:10s} rows {T}: xy wavefronts/row {a / T:.2f} (ideal 4), z {b / T:.2f} (ideal 2), total {(a + b) / T:.2f} (ideal 6)")
print("greedy overflow entries", tot_overflow, leftover)
