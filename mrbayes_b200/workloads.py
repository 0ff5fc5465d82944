"""Synthetic workloads for tests and bench.py: random unrooted trees in MrBayes' node
numbering, reversible substitution models with their eigensystems, discrete-Gamma rates,
random alignments as state-set masks, and a per-chain buffer-index allocator that plays the
role of MrBayes' condLikeIndex / tiProbsIndex / nodeScalerIndex tables (with a private
scratch slot per chain, so that several chains can be in flight in one launch).

No likelihood arithmetic lives here -- only inputs.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import abi


# ----------------------------------------------------------------------------- models
def discrete_gamma_rates(alpha: float, K: int) -> np.ndarray:
    """Mean rate of each of K equal-probability Gamma(alpha, alpha) categories
    (what DiscreteGamma, reference src/utils.c:10500, returns with median=0)."""
    if K == 1:
        return np.ones(1)
    from scipy.special import gammainc
    from scipy.stats import gamma as gdist
    cuts = gdist.ppf(np.arange(1, K) / K, a=alpha, scale=1.0 / alpha)
    edges = np.concatenate([[0.0], cuts, [np.inf]])
    # E[x; a<x<b] for Gamma(alpha, rate alpha) = P(alpha+1, alpha*b) - P(alpha+1, alpha*a)
    cdf1 = np.where(np.isinf(edges), 1.0, gammainc(alpha + 1.0, alpha * np.where(np.isinf(edges), 1.0, edges)))
    return (cdf1[1:] - cdf1[:-1]) * K


def reversible_model(S: int, rng: np.random.Generator, freqs=None, exch=None):
    """Random GTR-like reversible model scaled to one expected substitution per unit time.
    -> (freqs[S], V[S,S], Vinv[S,S], lambda[S]) with Q = V diag(lambda) Vinv."""
    pi = rng.dirichlet(np.full(S, 5.0)) if freqs is None else np.asarray(freqs, float)
    if exch is None:
        R = rng.gamma(2.0, 1.0, size=(S, S))
        R = np.triu(R, 1)
        R = R + R.T
    else:
        R = np.asarray(exch, float)
    Q = R * pi[None, :]
    np.fill_diagonal(Q, 0.0)
    np.fill_diagonal(Q, -Q.sum(1))
    Q /= -(pi * np.diag(Q)).sum()
    # symmetrise: B = D^1/2 Q D^-1/2 is symmetric for reversible Q
    d = np.sqrt(pi)
    B = (d[:, None] * Q) / d[None, :]
    B = 0.5 * (B + B.T)
    lam, U = np.linalg.eigh(B)
    V = U / d[:, None]
    Vinv = U.T * d[None, :]
    return pi, np.ascontiguousarray(V), np.ascontiguousarray(Vinv), lam


# ----------------------------------------------------------------------------- trees
@dataclass
class Tree:
    """Unrooted binary tree the way MrBayes stores it: tips 0..n-1, interior
    n..2n-3; ``root`` is a tip whose only neighbour ``root_left`` is the interior root with
    three neighbours (left, right and the root tip)."""
    n_tips: int
    left: np.ndarray       # [nNodes] child index or -1
    right: np.ndarray
    anc: np.ndarray        # [nNodes] ancestor or -1
    length: np.ndarray     # [nNodes] length of the branch above the node
    root: int              # the root tip
    post: list             # interior nodes in post-order (intDownPass)

    @property
    def n_nodes(self):
        return 2 * self.n_tips - 2

    @property
    def n_int(self):
        return self.n_tips - 2

    @property
    def root_left(self):
        return self.post[-1]

    def path_to_root(self, node: int):
        """interior nodes from anc(node) up to the interior root, in post-order."""
        out = []
        p = self.anc[node]
        while p >= 0 and p != self.root:
            out.append(int(p))
            p = self.anc[p]
        return out


def random_tree(n_tips: int, rng: np.random.Generator, mean_len: float = 0.1) -> Tree:
    assert n_tips >= 3
    n_nodes = 2 * n_tips - 2
    left = np.full(n_nodes, -1)
    right = np.full(n_nodes, -1)
    anc = np.full(n_nodes, -1)
    root = n_tips - 1                       # last tip hangs below the interior root
    nxt = n_tips                            # next interior index
    # start: interior root with tips 0, 1 as children and `root` as ancestor
    r = nxt; nxt += 1
    left[r], right[r], anc[r] = 0, 1, root
    anc[0] = anc[1] = r
    left[root] = r
    placed = [0, 1]
    nodes_with_branch = [0, 1]              # nodes whose upper branch can be split
    for tip in range(2, n_tips - 1):
        tgt = int(rng.choice(nodes_with_branch))
        a = int(anc[tgt])
        new = nxt; nxt += 1
        # splice `new` between tgt and its ancestor
        if left[a] == tgt:
            left[a] = new
        else:
            right[a] = new
        anc[new] = a
        if rng.random() < 0.5:
            left[new], right[new] = tgt, tip
        else:
            left[new], right[new] = tip, tgt
        anc[tgt] = new
        anc[tip] = new
        nodes_with_branch += [tip, new]
        placed.append(tip)
    assert nxt == n_nodes
    length = rng.exponential(mean_len, n_nodes)
    length[root] = 0.0
    post = []

    def visit(p):
        stack = [(p, 0)]
        while stack:
            node, st = stack.pop()
            if left[node] < 0 or node == root:
                continue
            if st == 0:
                stack.append((node, 1))
                stack.append((int(right[node]), 0))
                stack.append((int(left[node]), 0))
            else:
                post.append(node)
    visit(r)
    return Tree(n_tips, left, right, anc, length, root, post)


# ----------------------------------------------------------------------------- alignments
def random_masks(n_tips: int, C: int, S: int, rng: np.random.Generator, p_missing=0.02, p_ambig=0.0):
    """Tip observations as state-set masks [tip][C] (uint64).  Patterns are generated from a
    few 'ancestral' sequences so that sites are not pure noise (keeps lnL per site realistic)."""
    base = rng.integers(0, S, size=C)
    masks = np.zeros((n_tips, C), np.uint64)
    full = np.uint64((1 << S) - 1) if S < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    for t in range(n_tips):
        mut = rng.random(C) < 0.25
        st = np.where(mut, rng.integers(0, S, size=C), base)
        m = (np.uint64(1) << st.astype(np.uint64))
        miss = rng.random(C) < p_missing
        m = np.where(miss, full, m)
        if p_ambig > 0:
            amb = rng.random(C) < p_ambig
            extra = (np.uint64(1) << rng.integers(0, S, size=C).astype(np.uint64))
            m = np.where(amb, m | extra, m)
        masks[t] = m
    return masks


# ----------------------------------------------------------------------------- index tables
@dataclass
class ChainState:
    """Buffer indices of one chain: current and scratch slot per node (host-side flips)."""
    cl: np.ndarray          # [nNodes] current CL buffer (tips: the tip index)
    cl_scratch: np.ndarray
    ti: np.ndarray
    ti_scratch: np.ndarray
    ns: np.ndarray          # node scalers
    ns_scratch: np.ndarray
    site: int
    site_scratch: int
    scaled: np.ndarray      # node currently has a scaler
    eigen: int = 0


@dataclass
class Problem:
    """A division + chains on one instance, with evaluation builders."""
    S: int
    K: int
    C: int
    n_chains: int
    tree: list                      # one Tree per chain
    masks: np.ndarray
    weights: np.ndarray
    freqs: np.ndarray
    V: np.ndarray
    Vinv: np.ndarray
    lam: np.ndarray
    rates: np.ndarray
    p_invar: float = 0.0
    chains: list = field(default_factory=list)
    flags: int = 0

    # ---- geometry ------------------------------------------------------------------
    @property
    def n_tips(self):
        return self.tree[0].n_tips

    @property
    def n_nodes(self):
        return self.tree[0].n_nodes

    @property
    def n_int(self):
        return self.tree[0].n_int

    def config(self, max_evaluations=None):
        nch = self.n_chains
        return dict(tip_count=self.n_tips, partials_count=self.n_tips + 2 * nch * self.n_int,
                    state_count=self.S, pattern_count=self.C, category_count=self.K,
                    matrix_count=2 * nch * self.n_nodes, scaler_count=2 * nch * (self.n_int + 1),
                    eigen_count=max(nch, 1), weight_rows=1,
                    max_evaluations=max_evaluations or nch)

    def allocate(self):
        nT, nI, nN = self.n_tips, self.n_int, self.n_nodes
        self.chains = []
        for ch in range(self.n_chains):
            cl = np.arange(nN); cls = np.full(nN, -1)
            cl[nT:] = nT + (2 * ch) * nI + np.arange(nI)
            cls[nT:] = nT + (2 * ch + 1) * nI + np.arange(nI)
            ti = (2 * ch) * nN + np.arange(nN)
            tis = (2 * ch + 1) * nN + np.arange(nN)
            base = 2 * ch * (nI + 1)
            ns = np.full(nN, -1); nss = np.full(nN, -1)
            ns[nT:] = base + np.arange(nI)
            nss[nT:] = base + nI + 1 + np.arange(nI)
            self.chains.append(ChainState(cl, cls, ti, tis, ns, nss, base + nI, base + 2 * nI + 1,
                                          np.zeros(nN, bool), eigen=0))

    def create(self, lib: abi.Library, device=0, max_evaluations=None, flags=0) -> abi.Instance:
        inst = abi.Instance(lib, device=device, flags=flags, **self.config(max_evaluations))
        for t in range(self.n_tips):
            inst.set_tip_states(t, self.masks[t])
        inst.set_pattern_weights(0, self.weights)
        inst.set_eigen_decomposition(0, self.V, self.Vinv, self.lam)
        return inst

    # ---- evaluation builders ----------------------------------------------------------
    def _spec(self, ch: int, dirty_branches, dirty_nodes, full: bool) -> abi.EvalSpec:
        """Mirror of LaunchLogLikeForDivision's index protocol (reference
        src/likelihood.c:7885-7967) on this chain's tables."""
        st, tr = self.chains[ch], self.tree[ch]
        # site scalers: flip, reset or copy
        st.site, st.site_scratch = st.site_scratch, st.site
        site_src = abi.NONE if full else st.site_scratch
        mats = []
        for node in dirty_branches:
            st.ti[node], st.ti_scratch[node] = st.ti_scratch[node], st.ti[node]
            mats.append((st.ti[node], st.eigen, tr.length[node]))
        ops = []
        for p in tr.post:
            if p not in dirty_nodes:
                continue
            st.cl[p], st.cl_scratch[p] = st.cl_scratch[p], st.cl[p]
            l, r = int(tr.left[p]), int(tr.right[p])
            is_root = (tr.anc[p] == tr.root)
            c3, m3 = (st.cl[tr.root], st.ti[p]) if is_root else (abi.NONE, abi.NONE)
            remove = st.ns[p] if (st.scaled[p] and not full) else abi.NONE
            st.ns[p], st.ns_scratch[p] = st.ns_scratch[p], st.ns[p]
            write = abi.NONE if is_root else st.ns[p]
            st.scaled[p] = not is_root
            ops.append((st.cl[p], st.cl[l], st.ti[l], st.cl[r], st.ti[r], c3, m3, write, remove))
        wk = (1.0 - self.p_invar) / self.K
        return abi.EvalSpec(
            mats=np.array(mats, abi.MAT_DTYPE) if mats else None,
            ops=np.array(ops, abi.OP_DTYPE) if ops else None,
            site_dst=st.site, site_src=site_src, root=int(st.cl[tr.root_left]), weights_row=0,
            flags=self.flags, p_invar=self.p_invar, has_p_invar=int(self.p_invar > 0),
            rates=self.rates / (1.0 - self.p_invar), cat_weights=np.full(self.K, wk), freqs=self.freqs, chain=ch)

    def full_evaluation(self, ch: int) -> abi.EvalSpec:
        tr = self.tree[ch]
        branches = [n for n in range(tr.n_nodes) if n != tr.root]
        return self._spec(ch, branches, set(tr.post), True)

    def branch_update(self, ch: int, node: int, new_length: float) -> abi.EvalSpec:
        """What a branch-length move leaves dirty: one P(t) and the path to the root.
        (The interior root's own branch matrix is always rebuilt, as in the reference.)"""
        tr = self.tree[ch]
        tr.length[node] = new_length
        branches = [node]
        if tr.root_left not in branches:
            branches.append(tr.root_left)
        if node == tr.root_left:
            dirty = {tr.root_left}
        else:
            dirty = set(tr.path_to_root(node))
        return self._spec(ch, branches, dirty, False)

    def random_branch_update(self, ch: int, rng: np.random.Generator) -> abi.EvalSpec:
        tr = self.tree[ch]
        node = int(rng.integers(0, tr.n_nodes))
        while node == tr.root:
            node = int(rng.integers(0, tr.n_nodes))
        return self.branch_update(ch, node, float(tr.length[node] * np.exp(0.5 * (rng.random() - 0.5))))

    def reject(self, ch: int, spec: abi.EvalSpec, old_lengths=None):
        """ResetFlips (reference src/mcmc.c:15695): undo the index swaps of `spec`."""
        st, tr = self.chains[ch], self.tree[ch]
        st.site, st.site_scratch = st.site_scratch, st.site
        inv_ti = {int(v): i for i, v in enumerate(st.ti)}
        for m in spec.mats:
            node = inv_ti[int(m["matrix"])]
            st.ti[node], st.ti_scratch[node] = st.ti_scratch[node], st.ti[node]
        inv_cl = {int(v): i for i, v in enumerate(st.cl)}
        for op in spec.ops:
            p = inv_cl[int(op["dest"])]
            st.cl[p], st.cl_scratch[p] = st.cl_scratch[p], st.cl[p]
            st.ns[p], st.ns_scratch[p] = st.ns_scratch[p], st.ns[p]
        if old_lengths is not None:
            tr.length[:] = old_lengths


def _copy_chain(st: ChainState) -> ChainState:
    return ChainState(st.cl.copy(), st.cl_scratch.copy(), st.ti.copy(), st.ti_scratch.copy(), st.ns.copy(),
                      st.ns_scratch.copy(), st.site, st.site_scratch, st.scaled.copy(), st.eigen)


def snapshot(pr: Problem):
    """Index tables and branch lengths of all chains (to return to later)."""
    return [_copy_chain(st) for st in pr.chains], [t.length.copy() for t in pr.tree]


def reset_evaluation(pr: Problem, ch: int, snap) -> abi.EvalSpec:
    """A full evaluation of chain `ch` that lands exactly on the snapshot's state: the tables
    are set to the snapshot with every slot pre-flipped, so the flips of the full evaluation
    write the snapshot's current slots.  (What a whole-tree move such as a rate-matrix update
    followed by acceptance looks like to the engine.)"""
    chains, lengths = snap
    st = _copy_chain(chains[ch])
    tr = pr.tree[ch]
    tr.length[:] = lengths[ch]
    st.site, st.site_scratch = st.site_scratch, st.site
    for n in range(tr.n_nodes):
        if n != tr.root:
            st.ti[n], st.ti_scratch[n] = st.ti_scratch[n], st.ti[n]
        if n >= tr.n_tips:
            st.cl[n], st.cl_scratch[n] = st.cl_scratch[n], st.cl[n]
            st.ns[n], st.ns_scratch[n] = st.ns_scratch[n], st.ns[n]
    pr.chains[ch] = st
    return pr.full_evaluation(ch)


def make_problem(S: int, K: int, C: int, n_tips: int, n_chains: int, seed: int, alpha=0.5,
                 p_invar=0.0, p_missing=0.02, p_ambig=0.0, same_tree=False, flags=None) -> Problem:
    rng = np.random.default_rng(seed)
    pi, V, Vinv, lam = reversible_model(S, rng)
    rates = discrete_gamma_rates(alpha, K)
    masks = random_masks(n_tips, C, S, rng, p_missing, p_ambig)
    weights = rng.integers(1, 4, size=C).astype(np.float32)
    t0 = random_tree(n_tips, rng)
    trees = [t0 if same_tree else random_tree(n_tips, rng) for _ in range(n_chains)]
    if flags is None:
        flags = abi.FLAG_NUC4_PINVAR_QUIRK if S == 4 else abi.FLAG_TIP_SHORTCUTS
    pr = Problem(S, K, C, n_chains, trees, masks, weights, pi, V, Vinv, lam, rates, p_invar, flags=flags)
    pr.allocate()
    return pr


# ----------------------------------------------------------------------------- variable-state (STANDARD data) divisions
@dataclass
class StdProblem(Problem):
    """Morphology-like division: every pattern has its own number of states (m->nStates), the
    equal-frequency Mk model, K gamma categories; tables as MrBayes lays them out for SYMPI_EQUAL
    (one [K][n][n] run of matrices per state count that occurs, in increasing n: tiIndex; one
    frequency vector per state count: bsIndex), `dummy` leading unobservable patterns
    (AddDummyChars, reference src/model.c:176-224) and the coding-bias correction at the root."""
    state_counts: np.ndarray = None
    matrix_offsets: np.ndarray = None
    freq_offsets: np.ndarray = None
    matrix_length: int = 0
    dummy: int = 0
    uncompressed: int = 0

    def config(self, max_evaluations=None):
        c = super().config(max_evaluations)
        c["flags"] = abi.CONFIG_VARIABLE_STATES
        return c

    def create(self, lib: abi.Library, device=0, max_evaluations=None, flags=0) -> abi.Instance:
        cfg = self.config(max_evaluations)
        cfg["flags"] |= flags
        inst = abi.Instance(lib, device=device, **cfg)
        inst.set_pattern_states(self.state_counts, self.matrix_offsets, self.freq_offsets, self.matrix_length,
                                self.dummy, self.uncompressed)
        for t in range(self.n_tips):
            inst.set_tip_states(t, self.masks[t])
        inst.set_pattern_weights(0, self.weights)
        return inst

    def _spec(self, ch, dirty_branches, dirty_nodes, full):
        sp = super()._spec(ch, dirty_branches, dirty_nodes, full)
        if len(sp.mats):
            sp.mats["eigen"] = abi.NONE
        return sp


def make_std_problem(C: int, K: int, n_tips: int, n_chains: int, seed: int, max_states=6, dummy=2,
                     alpha=0.8, p_missing=0.05, same_tree=False) -> StdProblem:
    rng = np.random.default_rng(seed)
    # a handful of state counts (the frequency table, one vector per count, must fit 64 entries)
    pool = sorted(set([2, 3, min(4, max_states), max(2, max_states // 2), max_states]))
    while sum(pool) > 64:
        pool.pop(-2)
    ns = rng.choice(np.array(pool), size=C).astype(np.int32)
    ns[:dummy] = 2
    if C > dummy:
        ns[dummy] = max_states                       # make sure the largest class occurs
    classes = sorted(set(int(n) for n in ns))
    ti_of, bs_of, ti, bs = {}, {}, 0, 0
    for n in classes:
        ti_of[n], bs_of[n] = ti, bs
        ti += n * n * K
        bs += n
    freqs = np.concatenate([np.full(n, 1.0 / n) for n in classes])
    masks = np.zeros((n_tips, C), np.uint64)
    base = rng.integers(0, 1 << 30, size=C)
    for t in range(n_tips):
        mut = rng.random(C) < 0.3
        st = np.where(mut, rng.integers(0, 1 << 30, size=C), base) % ns
        m = np.uint64(1) << st.astype(np.uint64)
        miss = rng.random(C) < p_missing
        m = np.where(miss, (np.uint64(1) << ns.astype(np.uint64)) - np.uint64(1), m)
        masks[t] = m
    # unobservable (dummy) patterns: all taxa in state 0, all taxa in state 1, ... (coding=variable)
    for d in range(dummy):
        masks[:, d] = np.uint64(1) << np.uint64(d % 2)
    weights = rng.integers(1, 4, size=C).astype(np.float32)
    weights[:dummy] = 0.0
    t0 = random_tree(n_tips, rng)
    trees = [t0 if same_tree else random_tree(n_tips, rng) for _ in range(n_chains)]
    rates = discrete_gamma_rates(alpha, K) * rng.uniform(0.5, 2.0)
    S = int(max(classes))
    pr = StdProblem(S, K, C, n_chains, trees, masks, weights, freqs, np.zeros((S, S)), np.zeros((S, S)), np.zeros(S),
                    rates, 0.0, flags=0,
                    state_counts=ns, matrix_offsets=np.array([ti_of[int(n)] for n in ns], np.int32),
                    freq_offsets=np.array([bs_of[int(n)] for n in ns], np.int32), matrix_length=ti,
                    dummy=dummy, uncompressed=int(weights.sum()))
    pr.allocate()
    return pr
