"""Multi-GPU sharding of the MS hot path over torch.distributed (one process per GPU).

Two modes:

* batch sharding -- stereo pairs are independent end to end: pair i goes to rank
  i % world, no collective on the data path (optionally an all_gather of the
  [N,H,W] disparities for reporting).  This is what bench.py scales.

* disparity-slab sharding -- for frames whose 32*D*H*W-byte volume does not fit one
  GPU (BASELINE config M: 1984x2880, D=640 -> 117 GB): rank r owns disparities
  [r*D/G, (r+1)*D/G) of the SAME pair.  Raw matcher costs at disparity d depend only on
  the two images, so each rank computes its slab alone; the AML channels need the
  per-pixel minimum and denominator over ALL disparities, which costs two small
  all-reduces ([4,h,w] floats each: min, then sum) between the three kernels phases.
  WTA merges through an order-preserving int64 key all-reduce(min); soft-argmin
  through an all_gather of per-pixel (max, sum e, sum d*e) partials.  The output stays
  sharded along D.  Cross-slab denominators are summed in a different order than the
  reference's sequential loop, so AML stays in its tolerance class (2e-6 abs), as
  SURVEY.md section 7 anticipates.

The collectives are torch.distributed calls (NCCL over NVLink on the GPU box, gloo in
the CPU tests of the host logic); the volumes never cross the links.
"""
import ctypes

from . import _lib


def shard_range(total, rank, world):
    """Contiguous block partition: (begin, count) of `total` items for `rank`."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    base, rem = divmod(int(total), world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def shard_batch(n_pairs, rank, world):
    """Round-robin pair -> rank assignment (pair i -> rank i % world)."""
    return list(range(rank, int(n_pairs), world))


def _dist():
    import torch.distributed as dist
    return dist


# ------------------------------------------------------------------ merges --
def merge_min(t, group=None):
    """in-place all-reduce(min) of the per-pixel slab minima [.., 4, h, w]."""
    dist = _dist()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def merge_sum(t, group=None):
    """in-place all-reduce(sum) of the partial AML denominators [.., 4, h, w]."""
    dist = _dist()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def merge_wta_keys(keys, group=None):
    """in-place all-reduce(min) of int64 WTA keys (cost bits made monotonic << 32 | d,
    top bit flipped so the signed order equals the unsigned one): the global winner and,
    on ties, the lowest disparity -- np.argmin's rule (main_msnet.py:444-448)."""
    dist = _dist()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


def gather_parts(part, group=None):
    """all_gather of per-rank soft-argmin partials [N,3,H,W] -> [world,N,3,H,W]."""
    import torch
    dist = _dist()
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return part.unsqueeze(0)
    world = dist.get_world_size(group)
    part = part.contiguous()
    out = torch.empty((world * part.shape[0],) + tuple(part.shape[1:]), dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(out, part, group=group)   # concatenates along dim 0
    return out.view((world,) + tuple(part.shape))


def wta_key_pack(cost_min, d_index):
    """Host/torch restatement of the device key (fte.cu: pack_key) for CPU tests and for
    callers that already hold (min, argmin): int64 tensor."""
    import torch
    bits = cost_min.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    neg = (bits & 0x80000000) != 0
    mono = torch.where(neg, (~bits) & 0xFFFFFFFF, bits | 0x80000000)
    u = (mono << 32) | (d_index.to(torch.int64) & 0xFFFFFFFF)
    return u ^ (-0x8000000000000000)


def wta_key_unpack(keys):
    """-> (argmin int32, min float32)."""
    import torch
    u = keys ^ (-0x8000000000000000)
    d = (u & 0xFFFFFFFF).to(torch.int32)
    mono = (u >> 32) & 0xFFFFFFFF
    pos = (mono & 0x80000000) != 0
    bits = torch.where(pos, mono & 0x7FFFFFFF, (~mono) & 0xFFFFFFFF)
    bits = torch.where(bits >= 0x80000000, bits - 0x100000000, bits).to(torch.int32)
    return d, bits.view(torch.float32)


def softargmin_merge_reference(parts):
    """torch restatement of msn_soft_argmin_merge_dev for CPU tests: parts [P,N,3,H,W]."""
    import torch
    m, s, t = parts[:, :, 0], parts[:, :, 1], parts[:, :, 2]
    M = m.max(dim=0).values
    sc = torch.exp(m - M.unsqueeze(0))
    return (t * sc).sum(0) / (s * sc).sum(0)


# ---------------------------------------------------------- slab extractor --
class SlabShardedMSFeatures(object):
    """Disparity-slab sharded MS volume: every rank calls this with the SAME pair(s) and
    receives its [N,8,D/G,h,w] slab.  Three kernel phases with two all-reduces between
    them (see module docstring)."""

    def __init__(self, N, H, W, maxdisp=192, rank=None, world=None, group=None, device=None, **kw):
        import torch

        from . import cbmv
        dist = _dist()
        if not torch.cuda.is_available():
            raise _lib.MsnetsError("SlabShardedMSFeatures needs a CUDA device (no CPU fallback)")
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.torch, self.group, self.rank, self.world = torch, group, rank, world
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.N, self.H, self.W = int(N), int(H), int(W)
        self.d_begin, self.d_count = shard_range(maxdisp, rank, world)
        if self.d_count < 1:
            raise ValueError("more ranks than disparities")
        if not kw.get("left_only", True):
            raise NotImplementedError("slab sharding provides the 8-channel (left) volume")
        self.params = cbmv.make_params(maxdisp, d_begin=self.d_begin, d_count=self.d_count, **kw)
        self.shape = cbmv.output_shape(self.N, self.H, self.W, self.params)
        self.h, self.w = self.shape[3], self.shape[4]
        with torch.cuda.device(self.device):
            nbytes = _lib.lib().msn_ms_slab_workspace_bytes(self.N, self.H, self.W, ctypes.byref(self.params))
            if nbytes == 0:
                raise _lib.MsnetsError(_lib.lib().msn_last_error().decode())
            self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.mins = torch.empty((self.N, 4, self.h, self.w), dtype=torch.float32, device=self.device)
            self.den = torch.empty((self.N, 4, self.h, self.w), dtype=torch.float32, device=self.device)

    def __call__(self, left, right, out=None):
        torch, L = self.torch, _lib.lib()
        for t in (left, right):
            if t.dtype != torch.uint8 or tuple(t.shape) != (self.N, self.H, self.W) or not t.is_cuda \
                    or not t.is_contiguous():
                raise ValueError("expected contiguous uint8 CUDA tensors of shape %s" % ((self.N, self.H, self.W),))
        if out is None:
            out = torch.empty(self.shape, dtype=torch.float32, device=self.device)
        p = ctypes.byref(self.params)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.msn_ms_slab_phase_a_dev(left.data_ptr(), right.data_ptr(), self.N, self.H, self.W, p, None,
                                                 out.data_ptr(), self.mins.data_ptr(), self.workspace.data_ptr(),
                                                 self.workspace.numel(), st))
            merge_min(self.mins, self.group)
            _lib.check(L.msn_ms_slab_phase_b_dev(out.data_ptr(), self.mins.data_ptr(), self.N, self.h, self.w, p,
                                                 self.den.data_ptr(), st))
            merge_sum(self.den, self.group)
            _lib.check(L.msn_ms_slab_phase_c_dev(out.data_ptr(), self.mins.data_ptr(), self.den.data_ptr(), self.N,
                                                 self.h, self.w, p, st))
        return out


class ExchangeSlabMSFeatures(object):
    """Disparity-slab sharded MS volume with the exchange fused INTO the kernel (msn_ms_slab_fused_dev):
    every rank calls this with the SAME pair(s) and receives its [N,8,D/G,h,w] slab.  A tile's raw costs
    never leave shared memory; the per-pixel AML minima and partial denominators travel between the
    ranks' GPUs inside the kernel, through exchange tables in peer-mapped device memory (NVLink /
    NVSwitch writes), so the volume is written exactly once and torch.distributed only carries the
    64-byte CUDA IPC handles at construction.

    One process per GPU: the tables are cudaMalloc allocations exported with cudaIpcGetMemHandle and
    all-gathered over `group`.  `connect=False` leaves the wiring to the caller (`table_ptr`, `wire`):
    several virtual ranks inside one process, as the one-GPU tests do."""

    # sub-slab widths tried in turn.  192 is the widest a tile parks on the TMA path and the most efficient
    # (per-tile fixed costs amortise over more disparities: 0.94 ms per config-B pair equivalent at 192, 1.07 at
    # 128, 1.32 at 80 -- profiles/xchg_grid.py)
    TILE_D_CHOICES = (192,)

    @staticmethod
    def sub_slabs(maxdisp, world):
        """Sub-slabs per rank: every rank's slab cut into the same number of equal parts, as narrow as the
        first width of TILE_D_CHOICES that divides every slab evenly; None when none does."""
        counts = [shard_range(maxdisp, r, world)[1] for r in range(world)]
        if min(counts) < 1:
            return None
        for width in ExchangeSlabMSFeatures.TILE_D_CHOICES:
            subs = -(-max(counts) // width)
            if subs * world <= 32 and all(c % subs == 0 for c in counts):
                return subs
        return None

    @staticmethod
    def default_row_bands(maxdisp, world, min_slab=320):
        """Row bands for `world` ranks on a `maxdisp`-disparity frame: a tile amortises its fixed costs over the
        disparities it holds (0.94 ms per config-B pair equivalent at 192 per CTA, 1.32 at 80), so slabs stay at
        least `min_slab` wide and the remaining ranks split the ROWS, which costs no exchange at all.  Measured on
        config M with 8 GPUs: 8 slabs 10.7 ms, 4 slabs x 2 bands 8.4, 2 x 4 8.2, 1 x 8 8.3."""
        slabs = max(1, min(world, int(maxdisp) // int(min_slab)))
        while world % slabs:
            slabs -= 1
        return world // slabs

    def __init__(self, N, H, W, maxdisp=192, rank=None, world=None, group=None, device=None, connect=True,
                 row_bands=1, **kw):
        """row_bands > 1: 2-D sharding of one frame -- rank = band * slabs + slab with slabs = world / row_bands:
        the rank computes disparity slab `slab` of row band `band` of the cropped image; only the `slabs` ranks of a
        band trade minima / denominators (row bands need no exchange).  Output [N, 8, D / slabs, h / row_bands, w]."""
        import torch

        from . import cbmv
        dist = _dist()
        if not torch.cuda.is_available():
            raise _lib.MsnetsError("ExchangeSlabMSFeatures needs a CUDA device (no CPU fallback)")
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        row_bands = int(row_bands)
        if row_bands < 1 or world % row_bands:
            raise ValueError("row_bands must divide the number of ranks")
        slabs = world // row_bands
        if slabs > 8:
            raise ValueError("the fused slab exchange serves at most 8 slabs per row band (one NVSwitch node)")
        self.torch, self.group, self.rank, self.world = torch, group, rank, world
        self.row_bands, self.slabs, self.band, self.slab = row_bands, slabs, rank // slabs, rank % slabs
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.N, self.H, self.W = int(N), int(H), int(W)
        self.d_begin, self.d_count = shard_range(maxdisp, self.slab, slabs)
        if self.d_count < 1:
            raise ValueError("more slabs than disparities")
        if not kw.get("left_only", True):
            raise NotImplementedError("slab sharding provides the 8-channel (left) volume")
        self.subs = self.sub_slabs(maxdisp, slabs)
        if self.subs is None:
            raise ValueError("maxdisp=%d over %d slabs does not cut into equal sub-slabs: use SlabShardedMSFeatures"
                             % (maxdisp, slabs))
        h_full = self.H - 2 * int(kw.get("board_h", 10))
        self.row_begin, self.row_count = shard_range(h_full, self.band, row_bands) if row_bands > 1 else (0, 0)
        if row_bands > 1 and self.row_count < 1:
            raise ValueError("more row bands than rows")
        self.params = cbmv.make_params(maxdisp, d_begin=self.d_begin, d_count=self.d_count, row_begin=self.row_begin,
                                       row_count=self.row_count, **kw)
        self.shape = cbmv.output_shape(self.N, self.H, self.W, self.params)
        self.h, self.w = self.shape[3], self.shape[4]
        L = _lib.lib()
        self._table = ctypes.c_void_p()
        self._opened = []
        self.epoch = 0
        with torch.cuda.device(self.device):
            nbytes = L.msn_ms_slab_fused_workspace_bytes(self.N, self.H, self.W, ctypes.byref(self.params))
            self.table_bytes = L.msn_ms_slab_exchange_bytes(self.N, self.H, self.W, ctypes.byref(self.params), slabs,
                                                            self.subs)
            if nbytes == 0 or self.table_bytes == 0:
                raise _lib.MsnetsError(L.msn_last_error().decode())
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            _lib.check(L.msn_peer_alloc(self.table_bytes, ctypes.byref(self._table)))
        self.xchg = _lib.SlabExchange()
        self.xchg.world, self.xchg.rank, self.xchg.subs = slabs, self.slab, self.subs   # the exchange group = one band
        self.xchg.tables[self.slab] = self._table.value
        self.slab_group = None     # torch.distributed group of this band's ranks (row_bands > 1, set by _connect_ipc)
        if connect:
            self._connect_ipc()

    @property
    def table_ptr(self):
        return self._table.value

    def wire(self, table_ptrs):
        """table_ptrs[r] = device pointer (valid on this device) of rank r's exchange table, one per rank of the
        whole launch; only the pointers of this rank's row band are used."""
        if len(table_ptrs) != self.world or table_ptrs[self.rank] != self._table.value:
            raise ValueError("wire: expected one pointer per rank, own table at index rank")
        for s in range(self.slabs):
            self.xchg.tables[s] = table_ptrs[self.band * self.slabs + s]

    def _connect_ipc(self):
        torch, L, dist = self.torch, _lib.lib(), _dist()
        if self.world == 1:
            return
        if self.slabs == 1 and self.row_bands == 1:
            return
        handle = (ctypes.c_ubyte * 64)()
        with torch.cuda.device(self.device):
            _lib.check(L.msn_peer_export(self._table, handle))
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
            allh = torch.empty((self.world, 64), dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh.view(-1), mine, group=self.group)
            allh = allh.cpu().numpy()
            for s in range(self.slabs):
                r = self.band * self.slabs + s
                if r == self.rank:
                    continue
                buf = (ctypes.c_ubyte * 64)(*[int(v) for v in allh[r]])
                ptr = ctypes.c_void_p()
                _lib.check(L.msn_peer_open(buf, ctypes.byref(ptr)))
                self._opened.append(ptr)
                self.xchg.tables[s] = ptr.value
            if self.row_bands > 1:      # every rank creates every band's group (torch.distributed's rule)
                for b in range(self.row_bands):
                    grp = dist.new_group(list(range(b * self.slabs, (b + 1) * self.slabs)))
                    if b == self.band:
                        self.slab_group = grp
            else:
                self.slab_group = self.group
            dist.barrier(group=self.group)

    def empty_wta_parts(self):
        """Buffers for the `wta=` by-product: (argmin int32, min1, min2 float32), each [subs,N,4,h,w]."""
        torch, shp = self.torch, (self.subs, self.N, 4, self.h, self.w)
        return (torch.empty(shp, dtype=torch.int32, device=self.device),
                torch.empty(shp, dtype=torch.float32, device=self.device),
                torch.empty(shp, dtype=torch.float32, device=self.device))

    def __call__(self, left, right, out=None, wta=None):
        """wta: None or the tensors of empty_wta_parts(): this rank's WTA triples per sub-slab, to be merged
        over the ranks with slab_wta_merge()."""
        torch, L = self.torch, _lib.lib()
        for t in (left, right):
            if t.dtype != torch.uint8 or tuple(t.shape) != (self.N, self.H, self.W) or not t.is_cuda \
                    or not t.is_contiguous():
                raise ValueError("expected contiguous uint8 CUDA tensors of shape %s" % ((self.N, self.H, self.W),))
        if out is None:
            out = torch.empty(self.shape, dtype=torch.float32, device=self.device)
        wp = (None, None, None) if wta is None else tuple(t.data_ptr() for t in wta)
        self.epoch += 1
        self.xchg.epoch = self.epoch & 0xFFFFFFFF or 1
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(L.msn_ms_slab_fused_dev(left.data_ptr(), right.data_ptr(), self.N, self.H, self.W,
                                               ctypes.byref(self.params), ctypes.byref(self.xchg), out.data_ptr(),
                                               wp[0], wp[1], wp[2], self.workspace.data_ptr(),
                                               self.workspace.numel(), st))
        return out

    def close(self):
        L = _lib.lib()
        if self.torch.cuda.is_available():
            self.torch.cuda.synchronize(self.device)
        for ptr in self._opened:
            L.msn_peer_close(ptr)
        self._opened = []
        if self._table:
            L.msn_peer_free(self._table)
            self._table = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def wta_merge_reference(idx_parts, min1_parts, min2_parts):
    """torch restatement of msn_wta_merge_dev for the CPU tests: parts [P,...] ordered by ascending
    disparity range -> (argmin, min1, min2)."""
    import torch
    P = idx_parts.shape[0]
    a = torch.full_like(min1_parts[0], float("inf"))
    b = torch.full_like(a, float("inf"))
    k = torch.zeros_like(idx_parts[0])
    for p in range(P):
        v1, v2 = min1_parts[p], min2_parts[p]
        less = v1 < a
        b = torch.where(less, a, torch.where(v1 < b, v1, b))
        k = torch.where(less, idx_parts[p], k)
        a = torch.where(less, v1, a)
        b = torch.where(v2 < b, v2, b)
    return k, a, b


def gather_wta_parts(idx_parts, min1_parts, min2_parts, group=None):
    """all_gather of the per-rank WTA triples [subs, ...] -> [world*subs, ...], rank-major = ascending
    disparity ranges (the collective half of slab_wta_merge; gloo-testable)."""
    import torch
    dist = _dist()
    parts = [t.contiguous() for t in (idx_parts, min1_parts, min2_parts)]
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        gathered = []
        for t in parts:
            g = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            dist.all_gather_into_tensor(g, t, group=group)
            gathered.append(g)
        parts = gathered
    return parts


def _wta_merge_local(parts):
    import torch
    P = parts[0].shape[0]
    shp = tuple(parts[0].shape[1:])
    n = 1
    for v in shp:
        n *= v
    idx = torch.empty(shp, dtype=torch.int32, device=parts[0].device)
    m1 = torch.empty(shp, dtype=torch.float32, device=parts[0].device)
    m2 = torch.empty(shp, dtype=torch.float32, device=parts[0].device)
    with torch.cuda.device(parts[0].device):
        _lib.check(_lib.lib().msn_wta_merge_dev(parts[0].data_ptr(), parts[1].data_ptr(), parts[2].data_ptr(), P, n,
                                                idx.data_ptr(), m1.data_ptr(), m2.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    return idx, m1, m2


def slab_wta_merge(idx_parts, min1_parts, min2_parts, group=None):
    """Merge of the per-slab WTA triples (argmin with absolute disparity, smallest, second smallest) that
    the slab kernels emit as a by-product -- SURVEY.md 8e(3): all-gather the per-GPU pairs, reduce locally
    (msn_wta_merge_dev).  *_parts: this rank's [subs, ...] CUDA tensors; returns (argmin int32, min1, min2)
    over ALL disparities, on every rank.  This rank's sub-slabs are merged first, so one triple per rank
    crosses NVLink.  Peak ratio: confidence.pkrn_confidence(min1, min2)."""
    if not idx_parts.is_cuda:
        raise _lib.MsnetsError("slab_wta_merge: expected CUDA tensors (no CPU fallback)")
    parts = [t.contiguous() for t in (idx_parts, min1_parts, min2_parts)]
    dist = _dist()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if parts[0].shape[0] > 1:
            parts = [t.unsqueeze(0) for t in _wta_merge_local(parts)]
        parts = gather_wta_parts(parts[0], parts[1], parts[2], group)
    return _wta_merge_local(parts)


def slab_soft_argmin(logits_slab, d_begin, group=None):
    """Soft-argmin over a D-sharded logit volume: logits_slab [N,D/G,H,W] on each rank,
    d_begin = first disparity of the slab -> full-range disparity [N,H,W] on every rank."""
    import torch
    x = logits_slab.contiguous()
    N, Dn, H, W = x.shape
    part = torch.empty((N, 3, H, W), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    with torch.cuda.device(x.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(L.msn_soft_argmin_partial_dev(x.data_ptr(), N, Dn, H, W, int(d_begin), part.data_ptr(), st))
        parts = gather_parts(part, group)
        disp = torch.empty((N, H, W), dtype=torch.float32, device=x.device)
        _lib.check(L.msn_soft_argmin_merge_dev(parts.data_ptr(), parts.shape[0], N, H, W, disp.data_ptr(), st))
    return disp


def slab_wta(cost_slab, d_begin, layout="dhw", group=None):
    """Winner-take-all over a D-sharded cost volume -> (argmin int32, min float32) of the
    full disparity range on every rank."""
    import torch
    c = cost_slab.contiguous()
    lay = 0 if layout == "hwd" else 1
    D = c.shape[-1] if lay == 0 else c.shape[0]
    shp = tuple(c.shape[:-1]) if lay == 0 else tuple(c.shape[1:])
    n = 1
    for v in shp:
        n *= v
    keys = torch.empty(shp, dtype=torch.int64, device=c.device)
    L = _lib.lib()
    with torch.cuda.device(c.device):
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(L.msn_wta_keys_dev(c.data_ptr(), n, D, lay, int(d_begin), keys.data_ptr(), st))
        merge_wta_keys(keys, group)
        am = torch.empty(shp, dtype=torch.int32, device=c.device)
        m1 = torch.empty(shp, dtype=torch.float32, device=c.device)
        _lib.check(L.msn_wta_unpack_dev(keys.data_ptr(), n, am.data_ptr(), m1.data_ptr(), st))
    return am, m1
