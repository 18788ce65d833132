"""Synthetic single-cell count matrices and cell sets.

Counts ``X[g, c] ~ Poisson(mu_g * s_c)`` with ``mu_g = exp(Normal(-4, 2))`` and
``s_c = exp(Normal(0, 0.35))``: density ~0.08, ~4 % of genes detected in more
than half of the cells (non-zero median), ~1 % all-zero genes.

Two generators with the same distribution:
* ``counts_numpy``  - ``np.random.default_rng(seed)``, mu then s then Poisson in
  2000-gene chunks, exactly the recipe of section 8-d (tests, small configs);
* ``counts_torch``  - the same recipe drawn with a seeded torch generator on a
  device (bench-size matrices in seconds instead of minutes).

Cell sets:
* ``deobs_sets``    - region sizes log-uniform in [lo, hi], members from
  ``default_rng(1).choice(C, N, replace=False)``;
* ``derand_draws``  - the reference's call sequence, ``np.random.seed(0)`` then
  ``np.random.choice(np.arange(C), size=N, replace=False)`` per draw
  (``/root/reference/pySpade/DErand.py:27,197``).
"""
import numpy as np
import scipy.sparse as sp

GENE_CHUNK = 2000


def counts_numpy(G: int, C: int, seed: int = 0) -> sp.csr_matrix:
    rng = np.random.default_rng(seed)
    mu = np.exp(rng.normal(-4.0, 2.0, size=G))
    s = np.exp(rng.normal(0.0, 0.35, size=C))
    blocks = []
    for g0 in range(0, G, GENE_CHUNK):
        lam = mu[g0:g0 + GENE_CHUNK, None] * s[None, :]
        blocks.append(sp.csr_matrix(rng.poisson(lam).astype(np.int32)))
    out = sp.vstack(blocks, format="csr")
    out.sort_indices()
    return out


def counts_torch(G: int, C: int, seed: int = 0, device="cuda"):
    """Returns ``(indptr int64[G+1], indices int32[nnz], counts int32[nnz])`` as torch
    tensors on ``device`` (CSR, genes x cells, column indices ascending)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    mu = torch.exp(torch.randn(G, generator=gen, device=device, dtype=torch.float64) * 2.0 - 4.0)
    s = torch.exp(torch.randn(C, generator=gen, device=device, dtype=torch.float64) * 0.35)
    row_nnz, idx_parts, val_parts = [], [], []
    for g0 in range(0, G, GENE_CHUNK):
        lam = (mu[g0:g0 + GENE_CHUNK, None] * s[None, :]).to(torch.float32)
        x = torch.poisson(lam, generator=gen).to(torch.int32)
        nz = x != 0
        row_nnz.append(nz.sum(dim=1))
        pos = nz.nonzero(as_tuple=False)                 # row-major -> columns ascending per row
        idx_parts.append(pos[:, 1].to(torch.int32))
        val_parts.append(x[nz])
    row_nnz = torch.cat(row_nnz).to(torch.int64)
    indptr = torch.zeros(G + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(row_nnz, dim=0)
    return indptr, torch.cat(idx_parts), torch.cat(val_parts)


def deobs_sets(C: int, S: int, lo: int = 50, hi: int = 2000, seed: int = 1, uniform_int=False):
    rng = np.random.default_rng(seed)
    sets = []
    for _ in range(S):
        if uniform_int:
            N = int(rng.integers(lo, hi + 1))
        else:
            N = int(round(float(np.exp(rng.uniform(np.log(lo), np.log(hi))))))
        N = max(1, min(N, C))
        sets.append(rng.choice(C, N, replace=False).astype(np.int64))
    return sets


def derand_draws(C: int, N: int, S: int, seed: int = 0):
    """Reference RNG stream: legacy global RandomState, one ``choice`` per draw."""
    state = np.random.get_state()
    try:
        np.random.seed(seed)
        draws = [np.random.choice(np.arange(C), size=N, replace=False) for _ in range(S)]
    finally:
        np.random.set_state(state)
    return draws


def concat_sets(sets):
    """``(members int32[sum N], offsets int64[S+1])`` for the batched C-ABI entry."""
    offsets = np.zeros(len(sets) + 1, dtype=np.int64)
    if len(sets):
        offsets[1:] = np.cumsum([len(s) for s in sets])
    members = (np.concatenate([np.asarray(s, dtype=np.int32) for s in sets])
               if len(sets) else np.zeros(0, np.int32))
    return np.ascontiguousarray(members, dtype=np.int32), offsets


def sample_rows(indptr, indices, counts, G, C, n_genes):
    """Host float64 CPM rows of a density-stratified gene sample of a device-resident count
    matrix (torch CSR as returned by ``counts_torch``): genes evenly spaced over the order by
    number of expressing cells, values in the reference's CPM operation order
    (``utils.py:150-155``).  Returns ``(gene ids ascending, scipy CSR [len(ids), C])`` - what
    the CPU oracle is fed when the full matrix is too large for it."""
    import torch

    nnz_g = (indptr[1:] - indptr[:-1])
    order = torch.argsort(nnz_g, stable=True)
    pick = order[torch.linspace(0, G - 1, n_genes, device=order.device).round().long()]
    pick = torch.unique(pick)                      # sorted, unique gene ids
    lens = nnz_g[pick]
    starts = indptr[:-1][pick]
    sub_ptr = torch.zeros(pick.numel() + 1, dtype=torch.int64, device=pick.device)
    sub_ptr[1:] = torch.cumsum(lens, 0)
    pos = torch.arange(int(sub_ptr[-1]), device=pick.device)
    row = torch.repeat_interleave(torch.arange(pick.numel(), device=pick.device), lens)
    src = starts[row] + (pos - sub_ptr[:-1][row])
    colsum = torch.zeros(C, dtype=torch.int64, device=pick.device)
    colsum.index_add_(0, indices.long(), counts.long())
    sub_idx = indices[src].cpu().numpy()
    sub_cnt = counts[src].cpu().numpy()
    with np.errstate(divide="ignore"):
        inv = 1.0 / colsum.cpu().numpy()
    data = (sub_cnt.astype(np.float64) * 1e6) * inv[sub_idx]        # utils.py:150-155 order
    X = sp.csr_matrix((data, sub_idx, sub_ptr.cpu().numpy()), shape=(pick.numel(), C))
    return pick.cpu().numpy(), X
