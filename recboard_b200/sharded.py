"""Row-sharded item table across the GPUs of one box: rank r owns a contiguous block of
item rows (with its gradient and optimizer state); query rows are replicated.

The reference has no counterpart -- its only multi-GPU mode is DDP, which all-reduces the whole
dense (N,d) table gradient every step (``ddp_backend: nccl``, E4SRec/README.md:23).  Here the
exchange steps are tiny and explicit:

  CE forward   one all-gather of the per-rank (row_max, row_sumexp, label_logit)  [3*M floats/rank]
  CE backward  one all-reduce (SUM) of the partial dU (M,d), issued asynchronously in front of the dW pass (which
               does not depend on it); the dW shard is purely local
  top-K eval   one all-gather of the per-rank sorted (vals, ids) (B,K) + a K-way merge kernel

The collectives and the elementwise merge math below are device-agnostic torch (they are exercised
on CPU with gloo, world_size 2, in tests/test_sharded_gloo.py); the per-shard partials come from
the CUDA kernels through ``recboard_b200.ops`` -- there is no CPU compute path in the product.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of rank ``rank``: [start, end)."""
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def merge_rowstats(stats: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """stats (R,3,M) = per-rank (max, sumexp, label_logit) -> (global lse (M,), label_logit (M,))."""
    m, l, ll = stats[:, 0], stats[:, 1], stats[:, 2]
    mg = m.max(dim=0).values
    lg = (l * torch.exp(m - mg)).sum(dim=0)
    return mg + torch.log(lg), ll.sum(dim=0)


def allgather_rowstats(m: torch.Tensor, l: torch.Tensor, ll: torch.Tensor, group=None) -> torch.Tensor:
    base = getattr(m, "_base", None)
    if (base is not None and base.dim() == 2 and base.shape[0] == 3 and base.is_contiguous() and m.data_ptr() == base.data_ptr()
            and l.data_ptr() == base[1].data_ptr() and ll.data_ptr() == base[2].data_ptr()):
        local = base       # ops.ce_rowstats hands out the three rows of one (3,M) buffer: no stack copy
    else:
        local = torch.stack([m, l, ll]).contiguous()
    world = dist.get_world_size(group)
    out = torch.empty(world * local.shape[0], *local.shape[1:], dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)  # concatenates along dim 0 (gloo and nccl)
    return out.view(world, *local.shape)


def allgather_topk(vals: torch.Tensor, ids: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(B,K) float32 vals + int32 ids per rank -> (R,B,K) each, with ONE collective."""
    packed = torch.stack([vals.contiguous().view(torch.int32), ids.contiguous()]).contiguous()
    world = dist.get_world_size(group)
    out = torch.empty(world * packed.shape[0], *packed.shape[1:], dtype=torch.int32, device=packed.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    out = out.view(world, *packed.shape)
    return out[:, 0].contiguous().view(torch.float32), out[:, 1].contiguous()


class _ShardedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, U, W_full, labels, bias_shard, scale, row_start, group, precision, n_skip, accumulate):
        from . import ops
        W_shard = W_full[n_skip:] if n_skip else W_full
        ctx.fused_du = bool(ctx.needs_input_grad[0]) and ops.fused_du_supported(U, precision, scale)
        du = None
        if ctx.fused_du:  # the forward sweep also accumulates this shard's unnormalised dU
            m, l, ll, du = ops.ce_rowstats(U, W_shard, labels, bias_shard, scale, label_base=row_start,
                                           precision=precision, want_dU=True)
        else:
            m, l, ll = ops.ce_rowstats(U, W_shard, labels, bias_shard, scale, label_base=row_start, precision=precision)
        gathered = allgather_rowstats(m, l, ll, group)
        lse, llg = ops.rowstats_merge(gathered) if gathered.is_cuda else merge_rowstats(gathered)   # one launch on the GPU
        empty = torch.empty(0, device=U.device)
        ctx.save_for_backward(U, W_full, labels, bias_shard if bias_shard is not None else empty, lse,
                              du if du is not None else empty, m if du is not None else empty)
        ctx.has_bias = bias_shard is not None
        ctx.scale, ctx.row_start, ctx.group, ctx.precision, ctx.n_skip = scale, row_start, group, precision, n_skip
        ctx.leaf = W_full if (accumulate and W_full.is_leaf and W_full.requires_grad) else None
        return (lse - llg).mean()

    @staticmethod
    def backward(ctx, grad_out):
        from . import ops
        U, W_full, labels, bias, lse, du_un, row_max = ctx.saved_tensors
        bias = bias if ctx.has_bias else None
        need = ctx.needs_input_grad
        P = ctx.n_skip
        W_shard = W_full[P:] if P else W_full
        g = 1.0 / U.shape[0]
        gdev = grad_out.detach().float().reshape(1).contiguous()
        need_db = ctx.has_bias and need[3]
        dU = None
        if ctx.fused_du:
            dU = ops.ce_du_finish(du_un, row_max, lse, W_shard, labels, g, ctx.scale, ctx.row_start, gdev)
        # the partial dU travels while the dW pass (which does not depend on it) runs: the all-reduce is asynchronous
        work = None
        if dU is not None:
            work = dist.all_reduce(dU, op=dist.ReduceOp.SUM, group=ctx.group, async_op=True)
        dU2, dW, db = ops.table_gradient(U, W_full, P, labels, lse, g, bias, ctx.scale, ctx.row_start, ctx.precision, gdev,
                                         need[0] and not ctx.fused_du, need[1], need_db, ctx.leaf)
        if dU is None and dU2 is not None:
            dU = dU2
            work = dist.all_reduce(dU, op=dist.ReduceOp.SUM, group=ctx.group, async_op=True)
        if work is not None:
            work.wait()
        return (dU.to(U.dtype) if dU is not None else None, dW, None,
                db.to(bias.dtype) if db is not None else None, None, None, None, None, None, None)


def sharded_fused_ce(U: torch.Tensor, W_shard: torch.Tensor, labels: torch.Tensor, row_start: int,
                     bias_shard: Optional[torch.Tensor] = None, scale: float = 1.0, group=None,
                     precision: Optional[str] = None, n_skip: int = 0, accumulate: bool = False) -> torch.Tensor:
    """Full-catalog CE where this rank scores only rows [row_start, row_start+len(shard)) of the
    table.  ``labels`` are GLOBAL item ids; every rank returns the same loss; the table gradient is the
    local shard's (never all-reduced), ``U.grad`` is the full gradient.  ``n_skip`` / ``accumulate`` as in
    ``ops.fused_ce``: ``W_shard`` may be the rank's whole parameter (pad rows first) and its existing bf16 ``.grad``
    may receive dW inside the pass."""
    return _ShardedCE.apply(U, W_shard, labels, bias_shard, float(scale), int(row_start), group, precision, int(n_skip),
                            bool(accumulate))


@torch.no_grad()
def sharded_topk(U: torch.Tensor, W_shard: torch.Tensor, K: int, row_start: int,
                 seen_crow: Optional[torch.Tensor] = None, seen_col: Optional[torch.Tensor] = None,
                 bias_shard: Optional[torch.Tensor] = None, scale: float = 1.0, group=None,
                 precision: Optional[str] = None,
                 merge_fn: Optional[Callable] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Global masked top-K over a row-sharded table: local top-K with global ids, one all-gather,
    K-way merge.  The seen CSR (global ids) is replicated; each rank applies the part in its range."""
    from . import ops
    vals, ids = ops.topk_eval(U, W_shard, K, seen_crow, seen_col, bias=bias_shard, scale=scale,
                              id_base=row_start, precision=precision)
    if merge_fn is None and vals.is_cuda:   # one collective, merged where it lands: no packing / unpacking copies beyond the stack
        packed = torch.stack([vals.contiguous().view(torch.int32), ids.contiguous()])
        world = dist.get_world_size(group)
        out = torch.empty(world, *packed.shape, dtype=torch.int32, device=packed.device)
        dist.all_gather_into_tensor(out.view(world * 2, *vals.shape), packed, group=group)
        return ops.topk_merge_packed(out)
    av, ai = allgather_topk(vals, ids, group)
    return (merge_fn or ops.topk_merge)(av, ai)


# --------------------------------------------------------------------------------------
# input-side gather over a row-sharded table: self.Item.embeddings(seqs), SASRec/main.py:183
# --------------------------------------------------------------------------------------
class _ShardedGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table_shard, idx, row_start, padding_idx, group, accumulate):
        from . import ops
        # ids this rank does not own (and the global padding id) fall outside [0, n_shard) after the shift: rb_gather_rows
        # returns zero rows for them, so the SUM over ranks is the gathered tensor
        local = idx - row_start
        if padding_idx >= 0:
            local = torch.where(idx == padding_idx, torch.full_like(local, -1), local)
        out = ops.gather_rows_raw(table_shard.detach(), local.contiguous())
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        ctx.save_for_backward(local)
        ctx.shape, ctx.dtype = table_shard.shape, table_shard.dtype
        ctx.leaf = table_shard if (accumulate and table_shard.is_leaf and table_shard.requires_grad) else None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from . import ops
        # the gathered tensor is replicated, so is its gradient: every rank adds the rows of the ids IT owns (the
        # scatter-add skips ids outside its shard) -- no communication in the backward
        (local,) = ctx.saved_tensors
        go = grad_out.contiguous()
        if go.dtype not in (torch.float32, torch.bfloat16):
            go = go.float()
        go, flat = go.view(-1, ctx.shape[1]), local.view(-1)
        leaf = ctx.leaf
        if leaf is not None and leaf.grad is not None and leaf.grad.is_contiguous() and leaf.grad.shape == ctx.shape \
                and leaf.grad.dtype in (torch.float32, torch.bfloat16):
            ops.scatter_add_rows_(leaf.grad, go, flat, -1)
            return None, None, None, None, None, None
        gdt = ctx.dtype if ctx.dtype in (torch.float32, torch.bfloat16) else torch.float32
        g = torch.zeros(ctx.shape, dtype=gdt, device=grad_out.device)
        ops.scatter_add_rows_(g, go, flat, -1)
        return g.to(ctx.dtype), None, None, None, None, None


def sharded_gather_rows(table_shard: torch.Tensor, idx: torch.Tensor, row_start: int, padding_idx: int = -1, group=None,
                        accumulate: bool = False) -> torch.Tensor:
    """``table[idx]`` for GLOBAL row ids over a row-sharded table: every rank gathers the rows it owns (zeros
    elsewhere) and one all-reduce (SUM) assembles the replicated (..., d) result; rows with ``idx == padding_idx`` are
    zero.  The backward needs no communication: each rank scatter-adds the (replicated) upstream gradient of its own
    ids into its shard's gradient (``accumulate`` as in ``ops.gather_rows``)."""
    return _ShardedGather.apply(table_shard, idx, int(row_start), int(padding_idx), group, bool(accumulate))


# --------------------------------------------------------------------------------------
# the same gather over PEER MEMORY: every rank reads the rows where they live (its shard, or a peer's over NVLink)
# --------------------------------------------------------------------------------------
class PeerTable:
    """The row shards of one table on the GPUs of ONE box, mapped into every rank (CUDA IPC): ``gather(idx)`` reads
    each row from the shard that owns it -- local HBM or a peer's over NVLink -- in one kernel (``rb_gather_rows_peers``).
    No collective and no replicated activations travel: the all-reduce of ``sharded_gather_rows`` moves
    ``world x`` the bytes (52 MB at the bench shape, +0.6 ms per step on 2 GPUs) to deliver what ~50 MB x (world-1)/world
    of peer loads deliver here.

    Construction is collective (handles are exchanged with ``all_gather_object``).  Rank r's ``table_shard[n_skip:]``
    holds global ids ``[row_start_r, row_start_r + len)``; the ranges must tile one interval in rank order.  The mapping
    follows the parameter's STORAGE: rebuild the PeerTable if the parameter is re-allocated.  A rank that updates its
    shard (optimizer step) calls ``fence()`` before anyone gathers again."""

    def __init__(self, table_shard: torch.Tensor, row_start: int, n_skip: int = 0, group=None):
        import ctypes as C
        from . import _lib as L
        dev = L.require_cuda(table_shard)
        if not table_shard.is_contiguous():
            raise ValueError("PeerTable needs a contiguous shard")
        self.group, self.table, self.n_skip, self.row_start = group, table_shard, int(n_skip), int(row_start)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        body = table_shard.detach()[n_skip:]
        self.n_rows, self.d, self.dtype = body.shape[0], body.shape[1], body.dtype
        # Every rank takes part in both collectives whatever happens locally, and all ranks fail together: a mapping
        # that works on some ranks only would leave the others hanging in the next collective.
        mine, err = None, None
        try:
            handle = C.create_string_buffer(64)
            off = C.c_int64(0)
            L.call(dev, "rb_ipc_export", L.ptr(body), C.cast(handle, C.c_void_p), C.cast(C.pointer(off), C.c_void_p))
            mine = (bytes(handle.raw), int(off.value), self.row_start, self.n_rows, self.d, str(self.dtype))
        except RuntimeError as e:   # e.g. memory that CUDA IPC cannot export
            err = f"rank {self.rank}: {e}"
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._bases, ptrs, starts = [], [], []
        if err is None and any(e is None for e in everyone):
            err = "a peer could not export its shard"
        if err is None:
            try:
                for r, (h, o, rs, nr, d, dt) in enumerate(everyone):
                    if d != self.d or dt != str(self.dtype):
                        raise ValueError(f"rank {r} holds a ({nr},{d}) {dt} shard, this rank ({self.n_rows},{self.d}) {self.dtype}")
                    if starts and rs != starts[-1] + everyone[r - 1][3]:
                        raise ValueError("the ranks' row ranges must tile one interval in rank order")
                    starts.append(rs)
                    if r == self.rank:
                        ptrs.append(body.data_ptr())
                        continue
                    p_out, b_out = C.c_void_p(0), C.c_void_p(0)
                    hb = C.create_string_buffer(h, 64)
                    L.call(dev, "rb_ipc_open", C.cast(hb, C.c_void_p), o, C.cast(C.pointer(p_out), C.c_void_p),
                           C.cast(C.pointer(b_out), C.c_void_p))
                    self._bases.append(b_out.value)
                    ptrs.append(p_out.value)
                starts.append(starts[-1] + everyone[-1][3])
            except (RuntimeError, ValueError) as e:
                err = f"rank {self.rank}: {e}"
        ok = torch.tensor([0.0 if err else 1.0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if float(ok) == 0.0:
            for b_ in self._bases:
                L.call(dev, "rb_ipc_close", b_)
            self._bases = []
            raise RuntimeError("PeerTable: the shards could not be mapped on every rank"
                               + (f" ({err})" if err else " (another rank failed)"))
        self._ptrs = (C.c_void_p * self.world)(*ptrs)
        self._starts = (C.c_int64 * (self.world + 1))(*starts)
        self.first_id, self.end_id = starts[0], starts[-1]
        self._token = torch.zeros(1, device=dev)
        self.fence()

    def fence(self) -> None:
        """Order this rank's writes to its shard (optimizer step) before every peer's next gather: one 4-byte
        all-reduce on the compute stream."""
        dist.all_reduce(self._token, group=self.group)

    def close(self) -> None:
        from . import _lib as L
        self.fence()
        for b in self._bases:
            L.call(self.table.device, "rb_ipc_close", b)
        self._bases = []

    def gather(self, idx: torch.Tensor, padding_idx: int = -1, accumulate: bool = False) -> torch.Tensor:
        """``table[idx]`` for GLOBAL ids (ids outside the table's range -- e.g. ``padding_idx`` -- give zero rows).
        Backward: each rank scatter-adds the (replicated) upstream gradient of the ids IT owns into its shard's gradient;
        nothing is communicated."""
        return _PeerGather.apply(self.table, idx, self, int(padding_idx), bool(accumulate))


class _PeerGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table_shard, idx, peers, padding_idx, accumulate):
        from . import _lib as L
        dev = L.require_cuda(table_shard, idx)
        flat = idx.contiguous().view(-1).to(torch.int64)
        if padding_idx >= peers.first_id and padding_idx < peers.end_id:   # a padding id inside the table's range
            flat = torch.where(flat == padding_idx, torch.full_like(flat, peers.first_id - 1), flat)
        out = torch.empty(flat.numel(), peers.d, dtype=peers.dtype, device=dev)
        L.call(dev, "rb_gather_rows_peers", peers._ptrs, peers._starts, peers.world, L.ptr(flat), L.ptr(out), flat.numel(),
               peers.d, L.dtype_code(out), L.stream_ptr(dev))
        ctx.save_for_backward(flat)
        ctx.peers, ctx.shape, ctx.dtype = peers, table_shard.shape, table_shard.dtype
        ctx.leaf = table_shard if (accumulate and table_shard.is_leaf and table_shard.requires_grad) else None
        return out.view(*idx.shape, peers.d)

    @staticmethod
    def backward(ctx, grad_out):
        from . import ops
        (flat,) = ctx.saved_tensors
        p = ctx.peers
        # rows of this rank's shard: global id -> local row (the n_skip pad rows in front are never addressed);
        # foreign ids fall outside [n_skip, n_skip + n_rows) and are dropped by the scatter-add
        local = flat - (p.row_start - p.n_skip)
        local = torch.where((flat >= p.row_start) & (flat < p.row_start + p.n_rows), local, torch.full_like(local, -1))
        go = grad_out.contiguous()
        if go.dtype not in (torch.float32, torch.bfloat16):
            go = go.float()
        go = go.view(-1, ctx.shape[1])
        leaf = ctx.leaf
        if leaf is not None and leaf.grad is not None and leaf.grad.is_contiguous() and leaf.grad.shape == ctx.shape \
                and leaf.grad.dtype in (torch.float32, torch.bfloat16):
            ops.scatter_add_rows_(leaf.grad, go, local, -1)
            return None, None, None, None, None
        gdt = ctx.dtype if ctx.dtype in (torch.float32, torch.bfloat16) else torch.float32
        g = torch.zeros(ctx.shape, dtype=gdt, device=grad_out.device)
        ops.scatter_add_rows_(g, go, local, -1)
        return g.to(ctx.dtype), None, None, None, None


# --------------------------------------------------------------------------------------
# checkpoint contract: the shards (de)serialise to the reference's single state_dict key
# (``Item.embeddings.weight``, benchmark/Amazon2014Beauty_550_LOU/SASRec.json:236-252 records the module tree)
# --------------------------------------------------------------------------------------
def unshard_table(shard: torch.Tensor, n_rows: int, group=None) -> torch.Tensor:
    """All ranks' row blocks (``shard_bounds`` order, possibly ragged) -> the full (n_rows, d) table on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_rows, world, r) for r in range(world)]
    pad = max(b - a for a, b in sizes)
    mine = torch.zeros(pad, *shard.shape[1:], dtype=shard.dtype, device=shard.device)
    a, b = sizes[rank]
    if shard.shape[0] != b - a:
        raise ValueError(f"rank {rank} holds {shard.shape[0]} rows, shard_bounds says {b - a}")
    mine[:b - a] = shard.detach()
    out = torch.empty(world * pad, *shard.shape[1:], dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    out = out.view(world, pad, *shard.shape[1:])
    return torch.cat([out[r, :sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)


def sharded_state_dict(state_dict: dict, key: str, shard: torch.Tensor, n_rows: int, n_pads: int = 0,
                       pad_rows=None, group=None) -> dict:
    """A copy of ``state_dict`` whose ``key`` holds the reference's single (n_pads + n_rows, d) table assembled from
    the ranks' shards (every rank gets the same dict; save it on rank 0 as the reference's Coach does)."""
    full = unshard_table(shard, n_rows, group)
    if n_pads:
        head = pad_rows if pad_rows is not None else torch.zeros(n_pads, full.shape[1], dtype=full.dtype, device=full.device)
        full = torch.cat([head.to(full), full], dim=0)
    out = dict(state_dict)
    out[key] = full
    return out


def load_table_shard(state_dict: dict, key: str, world_size: int, rank: int, n_pads: int = 0) -> torch.Tensor:
    """This rank's row block of the reference's single-table checkpoint entry (pad rows dropped)."""
    full = state_dict[key][n_pads:]
    a, b = shard_bounds(full.shape[0], world_size, rank)
    return full[a:b].clone()
