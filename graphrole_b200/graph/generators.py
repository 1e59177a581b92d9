"""Synthetic benchmark graphs emitted directly as CSR arrays (torch ops on the target device).

BASELINE.json's configurations (Erdos-Renyi |V|=1M |E|=20M; Barabasi-Albert |V|=10M |E|=200M)
cannot go through NetworkX objects; these generators build the same families as arrays.
"""
import torch

from graphrole_b200.graph.csr import CSRGraph


def erdos_renyi_csr(n: int, m: int, seed: int = 0, device='cuda') -> CSRGraph:
    """G(n, m)-style undirected graph: m uniformly sampled node pairs, self loops dropped,
    repeated pairs collapsed (so |E| is marginally below m), both directions stored."""
    gen = torch.Generator(device=device).manual_seed(seed)
    src = torch.randint(0, n, (m,), device=device, generator=gen)
    dst = torch.randint(0, n, (m,), device=device, generator=gen)
    keep = src != dst
    return CSRGraph.from_edges_device(src[keep], dst[keep], n, directed=False)


def barabasi_albert_csr(n: int, m: int, seed: int = 0, device='cuda') -> CSRGraph:
    """Preferential-attachment (Barabasi-Albert) graph with m edges per new node.

    Node t >= m attaches m edges whose targets are drawn from the list of all previous edge
    endpoints (the classic "repeated nodes" formulation), i.e. proportionally to degree.
    The sequential draw is vectorised: a draw that lands on the *target* slot of an earlier
    edge is a pointer to that edge, and pointer chains are resolved by pointer jumping in
    O(log) rounds.  Multi-edges collapse when the CSR is built.
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    n_new = n - m                       # nodes m .. n-1 each add m edges
    e_total = n_new * m
    e = torch.arange(e_total, device=dev, dtype=torch.int64)
    b = torch.div(e, m, rounding_mode='floor')   # index of the source among new nodes
    # endpoints list before source b: 2*m entries per earlier source; the first source (b = 0)
    # attaches to the m seed nodes 0..m-1
    u = torch.rand(e_total, device=dev, dtype=torch.float64, generator=gen)
    pos = torch.clamp((u * (2 * m * b).to(torch.float64)).to(torch.int64),
                      max=torch.clamp(2 * m * b - 1, min=0))
    blk = torch.div(pos, 2 * m, rounding_mode='floor')
    off = pos - blk * 2 * m
    first = b == 0
    direct = (off >= m) | first
    val = torch.where(first, e % m, blk + m)          # resolved node id where direct
    val = torch.where(direct, val, torch.full_like(val, -1))
    ptr = torch.where(direct, e, blk * m + off)       # earlier edge whose target we copy
    del u, pos, blk, off, first, direct
    while True:
        unresolved = val < 0
        if not bool(unresolved.any()):
            break
        idx = unresolved.nonzero(as_tuple=True)[0]
        p = ptr[idx]
        pv = val[p]
        val[idx] = pv                                  # resolved if the pointee is
        ptr[idx] = torch.where(pv >= 0, p, ptr[p])     # else jump
    src = b + m
    return CSRGraph.from_edges_device(src, val, n, directed=False)
