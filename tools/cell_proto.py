"""Executable specification of the cell (nested-dissection-in-time) solve of the banded reduced system.

numpy model at tile granularity of what kernels_band.cu does on the device (same storage, same order of stages):
the time axis is cut into C cells  [Qa] A-> M <-B [Qb];  A runs forward, B backward (reversed storage), both leave
their Schur complement in the middle separator M; a chain that starts next to a boundary separator Q carries WB
dense "spike" tile rows Z (its coupling to Q, which fills in along the chain) and the Q x Q block FF.
Run:  python tools/cell_proto.py     (checks against numpy's dense solve for several cell counts)
"""
import numpy as np

T = 4  # tile edge of the model (32 on the device)


class Layout:
    def __init__(self, NT, WB, C):
        self.NT, self.WB, self.C = NT, WB, C
        # interior tiles to distribute: NT - C*WB (middles) - (C-1)*WB (boundaries), over 2C chains
        free = NT - (2*C - 1)*WB
        assert free >= 2*C*WB, "chains must be at least WB tiles long"
        base, rem = divmod(free, 2*C)
        lens = [base + (1 if k < rem else 0) for k in range(2*C)]
        self.cells = []
        t = 0
        for c in range(C):
            qa0 = t if c > 0 else -1
            if c > 0:
                t += WB
            a0 = t; t += lens[2*c]
            m0 = t; t += WB
            t += lens[2*c + 1]
            b1 = t
            self.cells.append(dict(qa0=qa0, a0=a0, m0=m0, b1=b1, has_qa=c > 0, has_qb=c < C - 1))
        assert t == NT


class Chain:
    """one band problem: columns [0, Kend) are factored, [Kend, NTloc) = near separator (receives the Schur complement)"""
    def __init__(self, NTloc, Kend, WB, spiked):
        self.NT, self.Kend, self.WB, self.spiked = NTloc, Kend, WB, spiked
        W1 = WB + 1
        self.tiles = np.zeros((NTloc, W1 + WB, T, T))   # [K][dd] band tile (K+dd, K); [K][W1+r] spike tile (Q tile r, K)
        self.FF = np.zeros((WB, WB, T, T))              # Q x Q block (lower block triangle used)
        self.rhs = np.zeros((NTloc, T))
        self.gF = np.zeros((WB, T))

    def band(self, i, j):      # local scalar positions i >= j
        I, J = i//T, j//T
        return self.tiles[J, I - J], i % T, j % T

    def factor(self):
        WB, W1, Kend = self.WB, self.WB + 1, self.Kend
        y = np.zeros_like(self.rhs)
        for K in range(self.NT):
            LKK = None
            for dd in range(0, min(WB, self.NT - 1 - K) + 1):
                I = K + dd
                acc = self.tiles[K, dd].copy()
                for J in range(max(0, I - WB), min(K, Kend)):
                    acc -= self.tiles[J, I - J] @ self.tiles[J, K - J].T
                if K < Kend:
                    if dd == 0:
                        acc = np.linalg.cholesky(acc); LKK = acc
                    else:
                        acc = np.linalg.solve(LKK, acc.T).T
                self.tiles[K, dd] = acc
            if self.spiked:
                for r in range(WB):
                    acc = self.tiles[K, W1 + r].copy()
                    for J in range(max(0, K - WB), min(K, Kend)):
                        acc -= self.tiles[J, W1 + r] @ self.tiles[J, K - J].T
                    if K < Kend:
                        acc = np.linalg.solve(LKK, acc.T).T
                    self.tiles[K, W1 + r] = acc
            v = self.rhs[K].copy()
            for J in range(max(0, K - WB), min(K, Kend)):
                v -= self.tiles[J, K - J] @ y[J]
            if K < Kend:
                y[K] = np.linalg.solve(LKK, v); self.rhs[K] = y[K]
            else:
                self.rhs[K] = v
        if self.spiked:
            for J in range(Kend):
                for r in range(WB):
                    self.gF[r] -= self.tiles[J, W1 + r] @ self.rhs[J]
                    for r2 in range(r + 1):
                        self.FF[r, r2] -= self.tiles[J, W1 + r] @ self.tiles[J, W1 + r2].T

    def backward(self, x_near, x_far):
        """x_near: [WB][T] solution of the near separator (local order); x_far: [WB][T] of the spike separator"""
        WB, W1, Kend = self.WB, self.WB + 1, self.Kend
        x = np.zeros((self.NT, T)); x[Kend:] = x_near
        for J in range(Kend - 1, -1, -1):
            v = self.rhs[J].copy()
            for I in range(J + 1, min(J + WB, self.NT - 1) + 1):
                v -= self.tiles[J, I - J].T @ x[I]
            if self.spiked:
                for r in range(WB):
                    v -= self.tiles[J, W1 + r].T @ x_far[r]
            x[J] = np.linalg.solve(self.tiles[J, 0].T, v)
        return x[:Kend]


def cell_solve(S, g, WB, C, rank=0, world=1, comm=None):
    """Solve S x = g by the cell scheme.  With world > 1 (comm = a torch.distributed-like object with reduce(tensor, dst)
    and all_reduce(tensor) on float64 tensors), S and g are THIS RANK'S PARTIAL sums of the reduced system: every rank
    places its contribution into the full layout, the buffers of cell c are summed at its owner c*world//C (a reduce per
    cell), owners factor their cells, the boundary system and the solution are all-reduced -- the exchange pattern of
    libdynoba's distributed reduced solve (api.cu: build_reduced / solve_step)."""
    n = S.shape[0]; NT = n//T; L = Layout(NT, WB, C); w = WB*T
    owner = lambda c: c*world//C
    mine = [c for c in range(C) if owner(c) == rank]
    chains = []
    for cell in L.cells:
        A = Chain(cell['m0'] - cell['a0'] + WB, cell['m0'] - cell['a0'], WB, cell['has_qa'])
        Bc = Chain(cell['b1'] - cell['m0'], cell['b1'] - cell['m0'] - WB, WB, cell['has_qb'])
        chains.append((A, Bc))

    def locate(i, j):
        """storage of entry (i >= j): returns (array, index tuple)"""
        ti, tj = i//T, j//T
        for c, cell in enumerate(L.cells):
            A, Bc = chains[c]
            qa0, a0, m0, b1 = cell['qa0']*T, cell['a0']*T, cell['m0']*T, cell['b1']*T
            if cell['has_qa'] and qa0 <= j < a0:
                if i < a0:      # Qa x Qa: FF of A, natural order
                    return A.FF, ((i - qa0)//T, (j - qa0)//T, (i - qa0) % T, (j - qa0) % T)
                assert i < m0   # spike of A: row = Q position j, column = chain position i
                return A.tiles, ((i - a0)//T, WB + 1 + (j - qa0)//T, (j - qa0) % T, (i - a0) % T)
            if a0 <= j < m0 + w and i < m0 + w:
                li, lj = i - a0, j - a0
                return A.tiles, (lj//T, li//T - lj//T, li % T, lj % T)
            if m0 <= j < b1 and m0 + w <= i < b1:
                li, lj = b1 - 1 - j, b1 - 1 - i      # reversed: row <-> column
                return Bc.tiles, (lj//T, li//T - lj//T, li % T, lj % T)
            if cell['has_qb'] and b1 <= i < b1 + w and m0 + w <= j < b1:
                s, lc = b1 + w - 1 - i, b1 - 1 - j   # spike of B: reversed separator position, reversed chain position
                return Bc.tiles, (lc//T, WB + 1 + s//T, s % T, lc % T)
        raise AssertionError((i, j))

    def rhs_locate(p):
        for c, cell in enumerate(L.cells):
            A, Bc = chains[c]
            qa0, a0, m0, b1 = cell['qa0']*T, cell['a0']*T, cell['m0']*T, cell['b1']*T
            if cell['has_qa'] and qa0 <= p < a0: return A.gF, ((p - qa0)//T, (p - qa0) % T)
            if a0 <= p < m0 + w: return A.rhs, ((p - a0)//T, (p - a0) % T)
            if m0 + w <= p < b1: return Bc.rhs, ((b1 - 1 - p)//T, (b1 - 1 - p) % T)
        raise AssertionError(p)

    bw = WB*T
    for j in range(n):
        for i in range(j, min(n, j + bw + 1)):
            if S[i, j] != 0.0:
                arr, ix = locate(i, j); arr[ix] += S[i, j]
    for p in range(n):
        arr, ix = rhs_locate(p); arr[ix] += g[p]

    # ---- multi-GPU exchange 1: a reduce per cell to its owner (tiles, Q x Q block of the A chain), rhs all-reduced
    if world > 1:
        import torch
        for c in range(C):
            A, Bc = chains[c]
            for arr in (A.tiles, Bc.tiles, A.FF):
                t = torch.from_numpy(arr.reshape(-1)); comm.reduce(t, dst=owner(c))
            for arr in (A.rhs, Bc.rhs, A.gF, Bc.gF):
                t = torch.from_numpy(arr.reshape(-1)); comm.all_reduce(t)

    # ---- stage 1: all chains of this rank's cells (one launch on the device)
    for c in mine:
        A, Bc = chains[c]
        A.factor(); Bc.factor()

    # ---- stage 2: cell separator systems [M | Qa | Qb], M eliminated
    def flip(blockmat):      # reverse a dense (k*T) matrix on both axes
        return blockmat[::-1, ::-1]
    def dense_lower(ch, K0):   # dense lower-triangular copy of the near-separator block of a chain
        D = np.zeros((w, w))
        for j in range(WB):
            for i in range(j, WB):
                D[i*T:(i + 1)*T, j*T:(j + 1)*T] = ch.tiles[K0 + j, i - j]
        return np.tril(D)
    def dense_spike(ch, K0):   # Q (rows) x near separator (cols)
        D = np.zeros((w, w))
        for j in range(WB):
            for r in range(WB):
                D[r*T:(r + 1)*T, j*T:(j + 1)*T] = ch.tiles[K0 + j, WB + 1 + r]
        return D
    def dense_ff(ch):
        D = np.zeros((w, w))
        for r in range(WB):
            for r2 in range(r + 1):
                D[r*T:(r + 1)*T, r2*T:(r2 + 1)*T] = ch.FF[r, r2]
        return np.tril(D)
    sym = lambda Lo: Lo + np.tril(Lo, -1).T
    cellsys = {}
    for c in mine:
        cell = L.cells[c]
        A, Bc = chains[c]
        MM = sym(dense_lower(A, A.Kend)) + flip(sym(dense_lower(Bc, Bc.Kend)))
        gM = A.rhs[A.Kend:].ravel() + Bc.rhs[Bc.Kend:].ravel()[::-1]
        LM = np.linalg.cholesky(MM)
        yM = np.linalg.solve(LM, gM)
        Wa = np.linalg.solve(LM, dense_spike(A, A.Kend).T).T if cell['has_qa'] else np.zeros((w, w))
        Wb = np.linalg.solve(LM, flip(dense_spike(Bc, Bc.Kend)).T).T if cell['has_qb'] else np.zeros((w, w))
        Saa = sym(dense_ff(A)) - Wa @ Wa.T
        Sbb = flip(sym(dense_ff(Bc))) - Wb @ Wb.T
        Sba = -Wb @ Wa.T
        ga = A.gF.ravel() - Wa @ yM
        gb = Bc.gF.ravel()[::-1] - Wb @ yM
        cellsys[c] = dict(LM=LM, yM=yM, Wa=Wa, Wb=Wb, Saa=Saa, Sbb=Sbb, Sba=Sba, ga=ga, gb=gb)

    # ---- stage 3: boundary-separator system (block tridiagonal over Q_0 .. Q_{C-2}): every rank adds the Schur
    # complements of ITS cells into a zeroed copy; multi-GPU exchange 2: all-reduce; then solved on every rank
    nq = C - 1
    xQ = np.zeros((max(nq, 0), w))
    if nq:
        G = np.zeros((nq*w, nq*w)); gq = np.zeros(nq*w)
        for c, cs in cellsys.items():
            if c > 0:
                G[(c - 1)*w:c*w, (c - 1)*w:c*w] += cs['Saa']; gq[(c - 1)*w:c*w] += cs['ga']
            if c < C - 1:
                G[c*w:(c + 1)*w, c*w:(c + 1)*w] += cs['Sbb']; gq[c*w:(c + 1)*w] += cs['gb']
            if 0 < c < C - 1:
                G[c*w:(c + 1)*w, (c - 1)*w:c*w] += cs['Sba']; G[(c - 1)*w:c*w, c*w:(c + 1)*w] += cs['Sba'].T
        if world > 1:
            import torch
            comm.all_reduce(torch.from_numpy(G.reshape(-1))); comm.all_reduce(torch.from_numpy(gq))
        xQ = np.linalg.solve(G, gq).reshape(nq, w)

    # ---- stage 4: back-substitution of this rank's cells; multi-GPU exchange 3: all-reduce of the zero-padded solution
    x = np.zeros(n)
    for c in mine:
        cell = L.cells[c]
        A, Bc = chains[c]; cs = cellsys[c]
        xa = xQ[c - 1] if cell['has_qa'] else np.zeros(w)
        xb = xQ[c] if cell['has_qb'] else np.zeros(w)
        xM = np.linalg.solve(cs['LM'].T, cs['yM'] - cs['Wa'].T @ xa - cs['Wb'].T @ xb)
        a0, m0, b1 = cell['a0']*T, cell['m0']*T, cell['b1']*T
        x[m0:m0 + w] = xM
        x[a0:m0] = A.backward(xM.reshape(WB, T), xa.reshape(WB, T)).ravel()
        x[m0 + w:b1] = Bc.backward(xM[::-1].reshape(WB, T), xb[::-1].reshape(WB, T)).ravel()[::-1]
        if cell['has_qb']:
            x[b1:b1 + w] = xb
    if world > 1:
        import torch
        comm.all_reduce(torch.from_numpy(x))
    return x


def random_band_spd(n, bw, rng):
    M = np.zeros((n, n))
    for i in range(n):
        for j in range(max(0, i - bw), i + 1):
            M[i, j] = rng.standard_normal()
    S = np.tril(M) @ np.tril(M).T          # band (bw) SPD
    S = np.where(np.abs(np.subtract.outer(np.arange(n), np.arange(n))) <= bw, S, 0.0)
    return S + n*np.eye(n)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    WB = 3
    for C in (1, 2, 3, 4):
        NT = (2*C - 1)*WB + 2*C*WB + 5 + C
        n = NT*T
        S = random_band_spd(n, WB*T - 1, rng); g = rng.standard_normal(n)
        x = cell_solve(S, g, WB, C)
        ref = np.linalg.solve(S, g)
        print(f"C={C} NT={NT} n={n}  max |x - ref| / max|ref| = {np.abs(x - ref).max()/np.abs(ref).max():.3e}")
        assert np.abs(x - ref).max() <= 1e-10*np.abs(ref).max()
    print("cell solve ok")
