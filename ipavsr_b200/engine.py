"""Device engine: executes a layer graph (ipavsr_b200.layers) forward and backward with the sm_100a kernels of
libipavsr_b200.so.  This is what `theano.function(...)` + `T.grad` are in the reference
(`runners/2stream_dct.py:263-279`): one call = forward (+ loss + full backward + update).

PyTorch is used for device memory (caching allocator), streams and torch.distributed only.  Every number is
produced by a kernel of this repository, called through the C-ABI with raw pointers on torch's current stream.
There is no CPU fallback: without a CUDA device or the shared library the engine raises.

Memory layout in HBM
  * activations: row-major float32 matrices of N*T rows (row = n*T + t) with the leading dimension padded to a
    multiple of 4 floats (16-byte rows for vector loads / TMA); a concat is never materialised — it is a list of
    column segments that the consuming GEMM walks as a K-split.
  * parameters: ONE flat float32 arena; each device tensor starts on a 256-float boundary.  Gradients and optimiser
    state are arenas of the same layout, so the update is one fused kernel over the arena and the data-parallel
    gradient all-reduce is one NCCL call.  LSTM gate matrices are stored stacked and gate-interleaved
    (column 4u+g), which makes the hoisted input projection a single GEMM and the recurrence loads float4.
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from . import layers as L
from .derived import Derived, DiffImages, DctFeatures

ACT = {'linear': 0, 'sigmoid': 1, 'rectify': 2, 'tanh': 3, 'leaky_rectify': 4, 'very_leaky_rectify': 5,
       'softplus': 6, 'elu': 7}
GEMM_MODES = {'fp32': 0, 'tf32x3': 1, 'tf32': 2, 'f16x3': 4}
OPT = {'adam': 0, 'adadelta': 1, 'sgd': 2, 'momentum': 3, 'nesterov': 4}
GATES = ('ingate', 'forgetgate', 'cell', 'outgate')
SEG = 256    # arena alignment in floats


def _ld8(c):
    """Leading dimension of a device matrix: a multiple of 8 floats, so that both the float32 rows (vector loads, TMA)
    and the fp16 hi/lo copies of the same shape (16-byte TMA strides) are aligned."""
    return (int(c) + 7) // 8 * 8


class DevMat(object):
    """A (rows x cols) float32 device matrix with leading dimension ld; `t` keeps the storage alive."""
    __slots__ = ('t', 'ptr', 'rows', 'cols', 'ld', 'chunks')

    def __init__(self, t, ptr, rows, cols, ld, chunks=None):
        self.t, self.ptr, self.rows, self.cols, self.ld = t, ptr, rows, cols, ld
        self.chunks = chunks     # [(row0, nrows, upload event)] for an input that is still arriving in row chunks

    def row_slice(self, r0, n):
        return DevMat(self.t, self.ptr + 4 * r0 * self.ld, n, self.cols, self.ld)

    def torch_view(self):
        off = (self.ptr - self.t.data_ptr()) // 4
        return self.t.view(-1)[off: off + (self.rows - 1) * self.ld + self.cols].as_strided(
            (self.rows, self.cols), (self.ld, 1)) if self.rows > 0 else self.t.new_zeros((0, self.cols))


_STAMP = [0]     # global modification counter shared by every arena (which copy of a shared parameter is the newest)


class ParamArena(object):
    """Flat device arena for every parameter of a network + the mapping Lasagne Param <-> device view."""

    @staticmethod
    def next_stamp():
        _STAMP[0] += 1
        return _STAMP[0]

    def stamp_of(self, p):
        return max(self.pstamp.get(p, 0), self.stamp_all)

    def touch_all(self):
        """Every parameter of this arena was just modified on the device (an optimiser step)."""
        self.stamp_all = self.next_stamp()

    def touch(self, params):
        st = self.next_stamp()
        for p in params:
            self.pstamp[p] = st

    def sync_from_peers(self):
        """Refresh every parameter another arena holds a more recently modified copy of (engines sharing parameters)."""
        if not self.shared:
            return
        for p in self.shared:
            mine = self.stamp_of(p)
            best, best_st = None, mine
            for a in p.arenas():
                if a is not self and a.stamp_of(p) > best_st:
                    best, best_st = a, a.stamp_of(p)
            if best is not None:
                self._view(self.flat, p).copy_(best._view(best.flat, p))
                self.pstamp[p] = best_st
                self.split_dirty = True

    def __init__(self, layer_list, device):
        import weakref
        self.pstamp, self.stamp_all, self.shared = {}, 0, []
        self.device = device
        self.tensors = {}      # key -> (offset, rows, cols, ld, trainable)
        self.bind = {}         # Param -> (key, index function)
        self.order = []
        off = 0
        aux = 0
        self.aux_tensors = {}

        def add(key, rows, cols, pad_ld=True):
            nonlocal off
            ld = _ld8(cols) if (pad_ld and rows > 1) else cols
            self.tensors[key] = (off, rows, cols, ld)
            self.order.append(key)
            off += (rows * ld + SEG - 1) // SEG * SEG

        def add_aux(key, n):
            nonlocal aux
            self.aux_tensors[key] = (aux, n)
            aux += (n + 3) // 4 * 4

        for l in layer_list:
            if isinstance(l, L.DenseLayer):
                add((l, 'W'), l.W.shape[0], l.W.shape[1])
                self.bind[l.W] = ((l, 'W'), None)
                if l.b is not None:
                    add((l, 'b'), 1, l.b.shape[0])
                    self.bind[l.b] = ((l, 'b'), None)
            elif isinstance(l, L.LSTMLayer):
                H, I = l.num_units, l.num_inputs
                add((l, 'W_in'), I, 4 * H)
                add((l, 'W_hid'), H, 4 * H, pad_ld=False)      # the recurrence kernels take a dense (H, 4H) matrix
                add((l, 'b'), 1, 4 * H)
                for g, name in enumerate(GATES):
                    self.bind[getattr(l, 'W_in_to_' + name)] = ((l, 'W_in'), ('gate', g))
                    self.bind[getattr(l, 'W_hid_to_' + name)] = ((l, 'W_hid'), ('gate', g))
                    self.bind[getattr(l, 'b_' + name)] = ((l, 'b'), ('gate', g))
                if l.peepholes:
                    add((l, 'peep'), 3, H, pad_ld=False)
                    self.bind[l.W_cell_to_ingate] = ((l, 'peep'), ('row', 0))
                    self.bind[l.W_cell_to_forgetgate] = ((l, 'peep'), ('row', 1))
                    self.bind[l.W_cell_to_outgate] = ((l, 'peep'), ('row', 2))
                add((l, 'cell_init'), 1, H)
                add((l, 'hid_init'), 1, H)
                self.bind[l.cell_init] = ((l, 'cell_init'), None)
                self.bind[l.hid_init] = ((l, 'hid_init'), None)
            elif isinstance(l, L.BatchNormLayer):
                F = l.beta.shape[0]
                add((l, 'beta'), 1, F)
                add((l, 'gamma'), 1, F)
                self.bind[l.beta] = ((l, 'beta'), None)
                self.bind[l.gamma] = ((l, 'gamma'), None)
                add_aux((l, 'mean'), F)
                add_aux((l, 'inv_std'), F)
                self.bind[l.mean] = ((l, 'mean'), 'aux')
                self.bind[l.inv_std] = ((l, 'inv_std'), 'aux')
            elif isinstance(l, L.AdaptiveElemwiseSumLayer):
                add((l, 'coeffs'), 1, len(l.coeffs))
                for i, c in enumerate(l.coeffs):
                    self.bind[c] = ((l, 'coeffs'), ('elem', i))
        self.n = off
        self.tail = off                 # [tail+0] = loss sum, [tail+1] = normaliser count
        total = off + SEG
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(total, dtype=torch.float32, device=device)
        self.aux = torch.zeros(max(aux, 4), dtype=torch.float32, device=device)
        self.state = {}
        self.flat_hi = self.flat_lo = None      # tf32 split of the parameters (3xTF32 mode), refreshed when dirty
        self.flat_h16 = self.flat_l16 = None    # fp16 hi/lo split (f16x3 mode) + per-tensor scale exponents
        self.exps = self.amaxs = self.seg_id = None
        seg = np.zeros(max(off // SEG, 1), dtype=np.int32)
        for i, k in enumerate(self.order):
            o, rows, cols, ld = self.tensors[k]
            seg[o // SEG: o // SEG + (rows * ld + SEG - 1) // SEG] = i
        self.seg_host = seg
        self.split_dirty = True
        # initial values: the newest copy of every parameter (the host master, or another engine's device copy when the
        # parameter is already bound — e.g. a function on an intermediate layer compiled after training), then bind
        params = list(self.bind.keys())
        stamp = self.next_stamp()
        for p in params:
            self._view(self.flat, p).copy_(torch.from_numpy(np.ascontiguousarray(p.get_value())).reshape(
                self._view(self.flat, p).shape))
            self.pstamp[p] = max([a.stamp_of(p) for a in p.arenas()] + [0])
        for p in params:
            peers = p.arenas()
            if peers:
                self.shared.append(p)
                for a in peers:
                    if p not in a.shared:
                        a.shared.append(p)
            p._bindings.append(weakref.ref(self))

    # -- views ------------------------------------------------------------------------------------------
    def mat(self, key, which='flat'):
        off, rows, cols, ld = self.tensors[key]
        buf = getattr(self, which) if isinstance(which, str) else which
        return DevMat(buf, buf.data_ptr() + 4 * off, rows, cols, ld)

    def aux_ptr(self, key):
        off, n = self.aux_tensors[key]
        return self.aux.data_ptr() + 4 * off

    def _view(self, buf, p):
        key, idx = self.bind[p]
        if idx == 'aux':
            off, n = self.aux_tensors[key]
            return self.aux[off: off + n]
        off, rows, cols, ld = self.tensors[key]
        m = buf[off: off + rows * ld].view(rows, ld)[:, :cols]
        if idx is None:
            if rows == 1:
                return m[0, 0] if p.shape == () else m[0].view(p.shape)
            return m
        kind, k = idx
        if kind == 'gate':
            return m[:, k::4] if rows > 1 or len(p.shape) == 2 else m[0, k::4]
        if kind == 'row':
            return m[k]
        if kind == 'elem':
            return m[0, k]
        raise KeyError(idx)

    def read(self, p, which='flat'):
        buf = getattr(self, which)
        return self._view(buf, p).detach().cpu().numpy().astype(np.float32).reshape(p.shape).copy()

    def write(self, p, value, stamp=None):
        self.split_dirty = True
        v = self._view(self.flat, p)
        v.copy_(torch.from_numpy(np.ascontiguousarray(value, dtype=np.float32)).reshape(v.shape))
        self.pstamp[p] = stamp if stamp is not None else self.next_stamp()

    def opt_state(self, name):
        if name not in self.state:
            self.state[name] = torch.zeros_like(self.flat)
        return self.state[name]

    def tensor_of(self, p):
        return self.bind[p][0]


class _PackPlan(object):
    """Index tables of the packed / length-sorted execution of one padded batch (csrc/pack.cu), built on the host
    from the utterance lengths and uploaded in one copy.

    The reference zero-pads every utterance to T (`utils/datagen.py:129-139`) and encodes all N*T rows
    (`modelzoo/pretrained_encoder.py:4-9`).  Here the encoder sees the M valid frames only, utterances sorted by
    decreasing length, plus ONE zero row (row M) whose output is the constant every padding frame takes;
    everything behind the encoder runs on the padded layout in the same sorted order, which makes the utterances of
    an LSTM tile equally long, so the recurrence kernels skip the frames at which their whole tile is masked.
      pack[r]   (M+1)  original padded row n*T+t of packed row r; -1 for the zero row
      valid[r]  (M)    sorted padded row of packed row r
      unpack[q] (N*T)  packed row of sorted padded row q (M for padding frames)
      perm[q]   (N*T)  original padded row of sorted padded row q;  unperm = its inverse
      order[i]  (N)    original utterance of sorted utterance i;    inv = its inverse
      mask      (N,T)  uint8, the mask in sorted order"""

    def __init__(self, lens, T):
        lens = np.asarray(lens, dtype=np.int64).reshape(-1)
        N = len(lens)
        self.N, self.T = N, int(T)
        order = np.argsort(-lens, kind='stable')
        ls = lens[order]
        off = np.zeros(N + 1, dtype=np.int64)
        off[1:] = np.cumsum(ls)
        M = int(off[N])
        self.M = M
        utt = np.repeat(np.arange(N, dtype=np.int64), ls)
        tt = np.arange(M, dtype=np.int64) - off[utt]
        pack = np.empty(M + 1, dtype=np.int32)
        pack[:M] = order[utt] * T + tt
        pack[M] = -1
        valid = (utt * T + tt).astype(np.int32)
        unpack = np.full(N * T, M, dtype=np.int32)
        unpack[valid] = np.arange(M, dtype=np.int32)
        perm = (order[:, None] * T + np.arange(T, dtype=np.int64)[None, :]).reshape(-1).astype(np.int32)
        unperm = np.empty(N * T, dtype=np.int32)
        unperm[perm] = np.arange(N * T, dtype=np.int32)
        inv = np.empty(N, dtype=np.int32)
        inv[order] = np.arange(N, dtype=np.int32)
        mask = (np.arange(T)[None, :] < ls[:, None]).astype(np.uint8).reshape(-1)
        mask4 = np.zeros((N * T + 3) // 4 * 4, dtype=np.uint8)
        mask4[:N * T] = mask
        self.order_host, self.perm_host, self.lens_sorted, self.offsets_host = order, perm, ls, off
        self.order32 = np.ascontiguousarray(order.astype(np.int32))
        # active_rows[t] = number of (sorted) utterances longer than t: the rows a per-step recurrent GEMM has to compute
        self.active_rows = np.ascontiguousarray((ls[None, :] > np.arange(T)[:, None]).sum(1).astype(np.int32))
        self.tables = [('pack', pack), ('valid', valid), ('unpack', unpack), ('perm', perm), ('unperm', unperm),
                       ('order', order.astype(np.int32)), ('inv', inv), ('mask', mask4.view(np.int32))]
        self.dev = None
        self.ready = None            # event: the tables are on the device

    def upload(self, device, pinned):
        """One host->device copy of all tables (through the pinned staging tensor `pinned`, int32, large enough)."""
        tot = sum(len(a) for _, a in self.tables)
        host = pinned[:tot]
        o = 0
        hn = host.numpy()
        for _, a in self.tables:
            hn[o:o + len(a)] = a
            o += len(a)
        self.dev = host.to(device, non_blocking=True)
        o = 0
        for name, a in self.tables:
            setattr(self, name, self.dev[o:o + len(a)])
            o += len(a)
        self.mask = self.mask.view(torch.uint8)[:self.N * self.T].view(self.N, self.T)
        self.offsets = torch.from_numpy(self.offsets_host).to(device, non_blocking=True)
        return tot


class _Profiler(object):
    """Optional per-entry-point device timing (IPAVSR_PROFILE=1): CUDA events around every C-ABI call, summed by
    entry-point name.  Off by default (it adds event overhead); used by tools/profile_step.py."""

    def __init__(self, by_shape=False):
        self.events = []
        self.by_shape = by_shape

    def call(self, name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(getattr(_lib.load(), name)(*args), name)
        e1.record()
        if self.by_shape and name.startswith('ipavsr_gemm'):
            # (transA, transB, M, N, K) lead the argument list of every GEMM entry point except ipavsr_gemm (mode first)
            a = args[1:6] if name == 'ipavsr_gemm' else args[0:5]
            name = '%s ta=%d tb=%d %dx%dx%d' % ((name,) + tuple(int(x) for x in a))
        self.events.append((name, e0, e1))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.events:
            n, t = out.get(name, (0, 0.0))
            out[name] = (n + 1, t + e0.elapsed_time(e1))
        self.events = []
        return out


class _Nvtx(object):
    """NVTX ranges around the phases of a step and around every layer (IPAVSR_NVTX=1): Nsight Systems / Compute group the
    kernels of the forward walk, the backward walk, the gradient all-reduce and the update by the reference layer names
    (`fc1_s1`, `lstm_s2`, `f_lstm_agg` ...).  Off by default: a push/pop pair per layer costs host time."""
    on = os.environ.get('IPAVSR_NVTX', '0') == '1'

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Nvtx.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *a):
        if _Nvtx.on:
            torch.cuda.nvtx.range_pop()


class _Run(object):
    """Per-call state: values, saved tensors for backward, gradients."""

    def __init__(self, N, T):
        self.N, self.T = N, T
        self.vals = {}
        self.saved = {}
        self.grads = {}
        self.plan = None        # _PackPlan when the batch runs packed / length-sorted
        self.packed = set()     # layers whose value holds the packed rows (M valid frames + the zero row)
        self.cat = {}           # ConcatLayer -> (materialised buffer, bound tensor)
        self.lstm_db_done = set()   # LSTMs whose bias gradient the backward kernel produced itself
        self.pending = {}       # layer -> CUDA event of work still running on a side stream
        self.bwd_ready = {}     # LSTM layer -> (event, dG) launched ahead of the backward walk
        self.keep = []          # buffers that must outlive the side-stream kernels
        self.xw_ready = {}      # LSTM layer -> input projection computed ahead of its turn (sibling LSTMs)
        self.use_branches = False   # small batch: every input branch on its own stream (Engine._branch_enter)
        self.branches_used = set()
        self.main_stream = None
        self.branch_joined = {}     # branch -> number of its layers the trunk has already waited for
        self.branch_done = {}       # branch -> number of its layers issued so far


class Engine(object):
    def __init__(self, output_layer, device=None, gemm_mode=None, lstm_impl=None, delta_exact=None, packed=None):
        if not torch.cuda.is_available():
            raise RuntimeError('ipavsr_b200 needs a CUDA device (B200, sm_100a); there is no CPU path')
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
        self.out = output_layer
        self.layers = L.get_all_layers(output_layer)
        # f16x3 (fp32-parity three-product arithmetic on the fp16 tensor cores + tensor-core LSTM recurrence) is the
        # shipped default; 'fp32' (CUDA-core FFMA everywhere) is the cross-check mode
        self.gemm_mode = GEMM_MODES[gemm_mode or os.environ.get('IPAVSR_GEMM_MODE', 'f16x3')]
        self.lstm_impl = int(os.environ.get('IPAVSR_LSTM_IMPL', '0')) if lstm_impl is None else int(lstm_impl)
        self.delta_exact = int(os.environ.get('IPAVSR_DELTA_EXACT', '1')) if delta_exact is None else int(delta_exact)
        with torch.cuda.device(self.device):
            self.arena = ParamArena(self.layers, self.device)
        self.requires_grad = {}
        for l in self.layers:
            ins = getattr(l, 'input_layers', None) or ([l.input_layer] if getattr(l, 'input_layer', None) else [])
            self.requires_grad[l] = bool(l.params) or any(self.requires_grad.get(i, False) for i in ins if i is not None)
        self.input_layers = [l for l in self.layers if isinstance(l, L.InputLayer)]
        self.mask_layers = set()
        for l in self.layers:
            if isinstance(l, L.LSTMLayer) and l.mask_incoming_index > 0:
                self.mask_layers.add(l.input_layers[1])
        # A ConcatLayer whose inputs are all LSTM outputs (late fusion) is materialised for free: every LSTM writes its
        # hidden states straight into its column slice of one buffer, so the consumers run ONE projection GEMM over the
        # whole width instead of a K-split of accumulating GEMMs (and one operand split instead of one per stream).
        consumers = {}
        for l in self.layers:
            for i in (getattr(l, 'input_layers', None) or [getattr(l, 'input_layer', None)]):
                if i is not None:
                    consumers.setdefault(i, []).append(l)
        # host inputs that only feed a DenseLayer (an encoder's fc1) can be uploaded in row chunks with the first GEMM
        # running chunk by chunk as they arrive (IPAVSR_CHUNKED_UPLOAD=1).  Off by default: measured, it does not help —
        # what is exposed is the total PCIe time of all streams, which function(...).prefetch hides instead.
        self.chunkable = set()
        for l in self.layers:
            if isinstance(l, L.InputLayer):
                c = l
                while consumers.get(c) and len(consumers[c]) == 1 and isinstance(consumers[c][0], L.ReshapeLayer):
                    c = consumers[c][0]
                cons = consumers.get(c, [])
                if (len(cons) == 1 and isinstance(cons[0], L.DenseLayer) and cons[0].nonlinearity.name != 'softmax' and
                        os.environ.get('IPAVSR_CHUNKED_UPLOAD', '0') == '1'):
                    self.chunkable.add(l)
        self.cat_plan, self.cat_of = {}, {}
        if os.environ.get('IPAVSR_MATERIALISE_CONCAT', '1') != '0':
            for l in self.layers:
                if (isinstance(l, L.ConcatLayer) and len(l.input_layers) > 1 and
                        all(isinstance(i, L.LSTMLayer) and consumers.get(i) == [l] for i in l.input_layers)):
                    offs, o = [], 0
                    for i in l.input_layers:
                        offs.append(o)
                        o += i.num_units
                    self.cat_plan[l] = (offs, o)
                    for i, off in zip(l.input_layers, offs):
                        self.cat_of[i] = (l, off)
        # Packed / length-sorted execution (_PackPlan): an input that only feeds a stack of (non-softmax) DenseLayers — a
        # DBNF encoder — is packed to its valid frames; the LAST layer of that stack expands back to the padded layout.
        # packed: None/'auto' = when a batch has >= 1/16 padding frames and >= 2048 rows, 'force' = whenever the mask is a
        # prefix mask (tests), 'off' = never (IPAVSR_PACKED=0).
        self.packed_mode = packed if packed is not None else {'0': 'off', '2': 'force'}.get(
            os.environ.get('IPAVSR_PACKED', '1'), 'auto')
        self.pack_in, self.pack_tail = {}, {}
        for l in self.input_layers:
            if l in self.mask_layers:
                continue
            chain, c = [], l
            while True:
                cons = consumers.get(c, [])
                if len(cons) != 1 or c is self.out:
                    break
                nxt = cons[0]
                if isinstance(nxt, L.ReshapeLayer):
                    c = nxt
                elif isinstance(nxt, L.DenseLayer) and nxt.nonlinearity.name != 'softmax':
                    chain.append(nxt)
                    c = nxt
                else:
                    break
            if chain:
                self.pack_in[l] = chain[-1]
                self.pack_tail[chain[-1]] = l
        # data parallel: the gradient all-reduce is issued in buckets while the backward walk is still running — the arena
        # is laid out in layer order and the walk finalises it from the tail (head, aggregate LSTMs, last stream ...) to
        # the front (first encoder), so every finished range [lo, hi) can go out while the encoders' weight gradients are
        # still being computed.  IPAVSR_AR_OVERLAP=0 restores the single all-reduce after the backward.
        self.ar_overlap = os.environ.get('IPAVSR_AR_OVERLAP', '1') != '0'
        self.ar_bucket_floats = int(os.environ.get('IPAVSR_AR_BUCKET', str(2 << 20)))
        self._ar_works, self._ar_hi = [], None
        self._layer_lo = {}
        for key, (off, rows, cols, ld) in self.arena.tensors.items():
            self._layer_lo[key[0]] = min(self._layer_lo.get(key[0], off), off)
        # pinned host streams of an encoder: 'dma' = one copy-engine transfer per utterance, 'gather' = the row-gather
        # kernel reading host memory (fewer driver calls, but its CTAs keep SMs from the compute kernels)
        self.host_upload = os.environ.get('IPAVSR_HOST_UPLOAD', 'dma')
        self._dct_basis = {}         # (image shape, K) -> DCT basis of a derived DCT stream
        self._plans = []             # most recent _PackPlans [(key, plan)]
        self._plan_pins = []         # ring of pinned staging tensors [(tensor, event)]
        self._lens_cache = {}
        self.dropout_seed = 1234
        self.dropout_calls = 0
        self.world = None            # (rank, world_size, group) when data-parallel
        self.step_t = np.float32(0)  # Adam's shared step counter (custom/updates.py:74)
        self._ws = None
        self._split_cache = {}
        self._split_by_storage = {}
        self._amax = {}
        self._lr_cache = None
        # independent LSTM recurrences (the per-stream LSTMs; the forward/backward aggregate pair) run concurrently
        # on side streams and overlap with the GEMMs of the other branches on the main stream
        self.concurrent_lstm = os.environ.get('IPAVSR_CONCURRENT_LSTM', '1') != '0'
        self._side = []
        self._side_next = 0
        self._copy_stream = None
        self._prefetched = []
        self._c14 = None
        self._st_pin = None
        # Branch streams.  A layer that depends on exactly one non-mask input (a stream's encoder, its DeltaLayer, its
        # LSTM) belongs to that input's branch; everything behind the fusion is the trunk.  Every branch runs its forward
        # and its backward on a stream of its own, ordered against the trunk by events only where data flows
        # (IPAVSR_BRANCH_STREAMS=0: one issue order on one stream; IPAVSR_BRANCH_ROWS: only for batches of at most that
        # many frames).  At the reference's own batch sizes (26 / 10 utterances) no kernel fills the machine and the single
        # issue order leaves the step 1.5x longer than its dependency chain (3.40 -> 2.36 ms, tools/timeline_small.py); at
        # 512 utterances the other branches' kernels fill the partial last waves of the tile-per-pair GEMMs, the hand-over
        # gaps between dependent kernels and the SMs a recurrence kernel leaves free (8.2 -> 7.5 ms).
        self.branch_mode = os.environ.get('IPAVSR_BRANCH_STREAMS', '1') != '0'
        self.branch_rows = int(os.environ.get('IPAVSR_BRANCH_ROWS', str(1 << 30)))
        from . import schedule
        self._branch_of, self._n_branches, self._trunk_fed = schedule.branch_assignment(self.layers, self.mask_layers)
        # sibling LSTMs: consecutive LSTM layers of the walk that read the same input (forward / backward direction of a
        # BLSTM).  Forward: their projections are all computed before the first recurrence starts.  Backward: they are
        # processed in the order their recurrences were launched (the walk would otherwise first wait for the one that
        # finishes LAST and only then issue the GEMMs of the one that finished first).
        self._lstm_siblings, self._bwd_order = schedule.lstm_sibling_groups(
            self.layers, self._branch_of, os.environ.get('IPAVSR_LSTM_SIBLINGS', '1') != '0')
        # data parallel: a bucket [lo, hi) of the gradient arena goes out once every layer at or behind lo has been walked.
        # Siblings are walked in forward order, i.e. the one with the LOWER arena offset first: the group's range is
        # released by its last member only
        self._flush_lo = dict(self._layer_lo)
        for l, others in self._lstm_siblings.items():
            grp = sorted((l,) + tuple(others), key=self.layers.index)
            if l is grp[-1]:
                los = [self._layer_lo[g] for g in grp if g in self._layer_lo]
                if los:
                    self._flush_lo[l] = min(los)
            else:
                self._flush_lo.pop(l, None)
        self._branch_streams = []
        self._cur_run = None
        # The loss is final long before the step is: right after the loss kernel (single process) or after the first
        # gradient bucket's all-reduce, which carries the arena's tail (data parallel).  It is copied to pinned host memory
        # at that point on a stream of its own, and the compiled function returns as soon as THAT copy has landed — with the
        # backward pass and the update still running.  The caller's next call then enqueues behind them, so the device
        # never waits for the host between steps (IPAVSR_EARLY_LOSS=0: read the loss after the update, as before).
        self.early_loss = os.environ.get('IPAVSR_EARLY_LOSS', '1') != '0'
        self._early_loss_ok = False          # set per call by function(...): no L2 term, not a graph replay
        self._loss_stream = self._loss_host = self._loss_ready = None
        # CUDA graphs for launch-bound small batches: 'auto' (default) | 'off'  (IPAVSR_GRAPH=0)
        self.graph_mode = 'off' if os.environ.get('IPAVSR_GRAPH', '1') == '0' else 'auto'
        self._graphs, self._graph_failed = {}, False
        self._zpool, self._zpos = None, 0
        self._opool, self._opos = None, 0
        # fp16 hi/lo of sigmoid/tanh outputs written by the GEMM epilogue itself (static scale 2^14, coalesced through the
        # epilogue's staging tile): saves the separate split pass over those activations (IPAVSR_EPILOGUE_SPLIT=0 disables)
        self.epilogue_split = os.environ.get('IPAVSR_EPILOGUE_SPLIT', '1') == '1'
        # DenseLayer backward prep emits dZ only as the fp16 operand pair (no float32 dZ, no max / split pass over it)
        self.fuse_prep = os.environ.get('IPAVSR_FUSE_PREP', '1') == '1'

    # ------------------------------------------------------------------------------------------------
    # small helpers
    # ------------------------------------------------------------------------------------------------
    @property
    def stream(self):
        # torch.cuda.current_stream costs ~7 us and a step asks ~120 times: the handle is pinned while a forward / backward
        # walk runs on one stream (the staging code, which switches streams, passes its handles explicitly)
        if self._st_pin is not None:
            return self._st_pin
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _pin_stream(self, on):
        self._st_pin = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream) if on else None

    def new(self, rows, cols, zero=False):
        ld = _ld8(cols)
        n = max(rows * ld, 4)
        t = (torch.zeros if zero else torch.empty)(n, dtype=torch.float32, device=self.device)
        return DevMat(t, t.data_ptr(), rows, cols, ld)

    def _side_stream(self):
        if not self._side:
            # the recurrence kernels run here: latency chains the main stream ends up waiting for, so their CTAs go first
            # whenever SMs free up (IPAVSR_SIDE_PRIORITY=0: same priority as the main stream)
            pr = -1 if os.environ.get('IPAVSR_SIDE_PRIORITY', '1') == '1' else 0
            self._side = [torch.cuda.Stream(device=self.device, priority=pr) for _ in range(3)]
        st = self._side[self._side_next % len(self._side)]
        self._side_next += 1
        return st

    def _fork(self):
        """Returns (side stream, its handle) ordered after everything queued so far on the main stream."""
        side = self._side_stream()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        side.wait_event(ev)
        return side, C.c_void_p(side.cuda_stream)

    def _wait(self, run, layer):
        ev = run.pending.pop(layer, None)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)

    def _join(self, run):
        for layer in list(run.pending.keys()):
            self._wait(run, layer)
        for layer, (ev, _) in list(run.bwd_ready.items()):
            torch.cuda.current_stream(self.device).wait_event(ev)

    # ---- branch streams (small batches) ------------------------------------------------------------------------------
    def _use_branches(self, N, T):
        return self.branch_mode and 2 <= self._n_branches <= 8 and N * T <= self.branch_rows

    def _branch_enter(self, run, b, sync_from_main):
        """Makes branch b's stream current (and the pinned kernel stream).  The first entry of a run — and any entry with
        sync_from_main — orders the branch behind everything queued on the trunk's stream so far."""
        while len(self._branch_streams) <= b:
            self._branch_streams.append(torch.cuda.Stream(device=self.device))
        bs = self._branch_streams[b]
        if sync_from_main or b not in run.branches_used:
            ev = torch.cuda.Event()
            ev.record(run.main_stream)
            bs.wait_event(ev)
            run.branches_used.add(b)
        torch.cuda.set_stream(bs)
        self._st_pin = C.c_void_p(bs.cuda_stream)

    def _branch_leave(self, run):
        torch.cuda.set_stream(run.main_stream)
        self._st_pin = C.c_void_p(run.main_stream.cuda_stream)

    def _branch_join(self, run, b):
        """The trunk waits for everything queued on branch b so far."""
        if b in run.branches_used:
            ev = torch.cuda.Event()
            ev.record(self._branch_streams[b])
            run.main_stream.wait_event(ev)

    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() * 4 < nbytes:
            self._ws = torch.empty((int(nbytes) + 3) // 4, dtype=torch.float32, device=self.device)
        return self._ws

    # ---- 3xTF32 operand splits: each tensor is split once per step and reused by every GEMM that reads it ----
    def _refresh_param_split(self):
        ar = self.arena
        if self.gemm_mode == 4:
            if ar.flat_h16 is None:
                ar.flat_h16 = torch.empty(ar.flat.numel(), dtype=torch.float16, device=self.device)
                ar.flat_l16 = torch.empty(ar.flat.numel(), dtype=torch.float16, device=self.device)
                ar.seg_id = torch.from_numpy(ar.seg_host).to(self.device)
                ar.exps = torch.zeros(len(ar.order) + 1, dtype=torch.int32, device=self.device)
                ar.amaxs = torch.zeros(len(ar.order) + 1, dtype=torch.float32, device=self.device)
            if ar.split_dirty and ar.n > 0:
                _lib.call('ipavsr_f16_split_segments', ar.flat.data_ptr(), ar.flat_h16.data_ptr(), ar.flat_l16.data_ptr(),
                          ar.n, ar.seg_id.data_ptr(), len(ar.order), ar.amaxs.data_ptr(), ar.exps.data_ptr(), self.stream)
                ar.split_dirty = False
            return
        if ar.flat_hi is None:
            ar.flat_hi = torch.empty_like(ar.flat)
            ar.flat_lo = torch.empty_like(ar.flat)
        if ar.split_dirty:
            _lib.call('ipavsr_tf32_split_rna', ar.flat.data_ptr(), ar.flat_hi.data_ptr(), ar.flat_lo.data_ptr(),
                      ar.flat.numel(), self.stream)
            ar.split_dirty = False

    def _split_of(self, m):
        """(hi, lo) DevMats of an operand: parameters come from the split arenas, activations are split on first use."""
        ar = self.arena
        if m.t is ar.flat:
            off = m.ptr - ar.flat.data_ptr()
            return (DevMat(ar.flat_hi, ar.flat_hi.data_ptr() + off, m.rows, m.cols, m.ld),
                    DevMat(ar.flat_lo, ar.flat_lo.data_ptr() + off, m.rows, m.cols, m.ld))
        key = (m.ptr, m.rows, m.cols, m.ld)
        hit = self._split_cache.get(key)
        if hit is None:
            n = (m.rows * m.ld + 3) // 4 * 4
            hi = torch.empty(n, dtype=torch.float32, device=self.device)
            lo = torch.empty(n, dtype=torch.float32, device=self.device)
            _lib.call('ipavsr_tf32_split_rna', m.ptr, hi.data_ptr(), lo.data_ptr(), n, self.stream)
            hit = (DevMat(hi, hi.data_ptr(), m.rows, m.cols, m.ld), DevMat(lo, lo.data_ptr(), m.rows, m.cols, m.ld), m.t)
            self._split_cache[key] = hit
        return hit[0], hit[1]

    def _zeros2(self):
        """Two zeroed floats ([|x|max, scale exponent] of an operand, a bound ...) carved out of one pre-zeroed pool per
        step: one fill kernel instead of one per request (66 of the 220 launches of a step were these fills)."""
        if self._zpool is None or self._zpos + 8 > self._zpool.numel():
            self._zpool = torch.zeros(2048, dtype=torch.float32, device=self.device)
            self._zpos = 0
        t = self._zpool[self._zpos: self._zpos + 2]
        self._zpos += 8                       # 32-byte slots
        return t

    def _ones2(self):
        """[1.0, 0]: the starting value of a bound max(1, ...) — from a pool filled once per step."""
        if self._opool is None or self._opos + 8 > self._opool.numel():
            self._opool = torch.zeros(32, 8, dtype=torch.float32, device=self.device)
            self._opool[:, 0] = 1.0
            self._opool = self._opool.view(-1)
            self._opos = 0
        t = self._opool[self._opos: self._opos + 2]
        self._opos += 8
        return t

    def _set_amax(self, m, t):
        """Registers the device max|m| produced alongside `m`.  The entry holds m's storage: while it exists the
        allocator cannot hand the same address to another buffer that would then read a stale 'ready' scale."""
        self._amax[(m.ptr, m.rows, m.cols, m.ld)] = (t, m.t)

    def _const14(self):
        """[amax = 1.0, exponent = 14] for tensors bounded by 1 (sigmoid / tanh activations, LSTM hidden states)."""
        if self._c14 is None:
            t = self._zeros2()
            t[0] = 1.0
            t[1:2].view(torch.int32)[0] = 14
            self._c14 = t
        return self._c14

    def _split16(self, m):
        """fp16x3 operand of a float32 DevMat: (hi ptr, lo ptr, exponent ptr); leading dimension = m.ld halves."""
        ar = self.arena
        if m.t is ar.flat:
            off = (m.ptr - ar.flat.data_ptr()) // 4
            return (ar.flat_h16.data_ptr() + 2 * off, ar.flat_l16.data_ptr() + 2 * off,
                    ar.exps.data_ptr() + 4 * int(ar.seg_host[off // SEG]))
        key = (m.ptr, m.rows, m.cols, m.ld)
        hit = self._split_cache.get(key)
        if hit is None and m.t is not None:
            # a row range of a tensor that is already split as a whole shares its hi/lo arrays and scale
            for pk in self._split_by_storage.get(id(m.t), ()):
                pptr, prows, pcols, pld = pk
                off = m.ptr - pptr
                if (pk in self._split_cache and pld == m.ld and pcols == m.cols and off >= 0 and off % (4 * pld) == 0 and
                        off // (4 * pld) + m.rows <= prows):
                    ph = self._split_cache[pk]
                    return ph[0].data_ptr() + off // 2, ph[1].data_ptr() + off // 2, ph[2].data_ptr() + 4
        if hit is None:
            self._split_by_storage.setdefault(id(m.t), []).append(key)
            n = max(m.rows * m.ld, 8)
            hi = torch.empty(n, dtype=torch.float16, device=self.device)
            lo = torch.empty(n, dtype=torch.float16, device=self.device)
            amax = self._amax.pop(key, None)
            if amax is not None:
                amax = amax[0]
            ready = amax is not None
            if amax is None:
                amax = torch.empty(2, dtype=torch.float32, device=self.device)
            _lib.call('ipavsr_f16_split', m.ptr, m.ld, m.rows, m.cols, hi.data_ptr(), lo.data_ptr(), m.ld,
                      amax.data_ptr(), amax.data_ptr() + 4, 1 if ready else 0, self.stream)
            hit = (hi, lo, amax, m.t)
            self._split_cache[key] = hit
        return hit[0].data_ptr(), hit[1].data_ptr(), hit[2].data_ptr() + 4

    def gemm(self, A, B, Cm, M, N, K, transA=0, transB=0, bias=None, act=0, accumulate=0, emit_split=False,
             amax_t=None):
        mode = self.gemm_mode
        if mode == 4:
            if not self.lib.ipavsr_gemm_f16_supported(M, N, K, 16, A.ld, 16, B.ld):
                mode = 0        # tiny / unaligned products: the exact FP32 kernel
            else:
                ah, al, ea = self._split16(A)
                bh, bl, eb = self._split16(B)
                amax = chi = clo = None
                if amax_t is not None:
                    amax = amax_t.data_ptr()        # the caller collects max|C| over several row-chunk products
                elif emit_split and not accumulate:
                    key = (Cm.ptr, Cm.rows, Cm.cols, Cm.ld)
                    if act in (1, 3) and self.epilogue_split:
                        # sigmoid / tanh outputs are bounded by 1: the epilogue writes the fp16 hi/lo split itself under
                        # the static scale 2^14, and no split pass is needed at all
                        n = max(Cm.rows * Cm.ld, 8)
                        hi = torch.empty(n, dtype=torch.float16, device=self.device)
                        lo = torch.empty(n, dtype=torch.float16, device=self.device)
                        self._split_cache[key] = (hi, lo, self._const14(), Cm.t)
                        chi, clo = hi.data_ptr(), lo.data_ptr()
                    else:
                        # the epilogue leaves max|C| behind, so the split of C (first use as an operand) needs no
                        # reduction pass
                        t = self._zeros2()
                        self._set_amax(Cm, t)
                        amax = t.data_ptr()
                _lib.call('ipavsr_gemm_f16x3', transA, transB, M, N, K, ah, al, A.ld, ea, bh, bl, B.ld, eb,
                          Cm.ptr, Cm.ld, bias, act, accumulate, amax, chi, clo, 14, self.stream)
                return
        if mode != 0 and not self.lib.ipavsr_gemm_tc_supported(transA, transB, M, N, K, A.ptr, A.ld, B.ptr, B.ld,
                                                                Cm.ptr, Cm.ld):
            mode = 0        # tiny / unaligned products: the exact FP32 kernel
        if mode == 1:
            ah, al = self._split_of(A)
            bh, bl = self._split_of(B)
            chi = clo = None
            if emit_split and not accumulate:
                n = (Cm.rows * Cm.ld + 3) // 4 * 4
                th = torch.empty(n, dtype=torch.float32, device=self.device)
                tl = torch.empty(n, dtype=torch.float32, device=self.device)
                chi, clo = th.data_ptr(), tl.data_ptr()
                self._split_cache[(Cm.ptr, Cm.rows, Cm.cols, Cm.ld)] = (
                    DevMat(th, chi, Cm.rows, Cm.cols, Cm.ld), DevMat(tl, clo, Cm.rows, Cm.cols, Cm.ld), Cm.t)
            _lib.call('ipavsr_gemm_tf32x3_presplit', transA, transB, M, N, K, ah.ptr, al.ptr, A.ld, bh.ptr, bl.ptr, B.ld,
                      Cm.ptr, Cm.ld, bias, act, accumulate, chi, clo, self.stream)
            return
        _lib.call('ipavsr_gemm', mode, transA, transB, M, N, K, A.ptr, A.ld, B.ptr, B.ld, Cm.ptr, Cm.ld,
                  bias, act, accumulate, None, 0, self.stream)

    def _proj(self, segs, W, out, bias, act, emit_split=False):
        """out = act( [segs...] @ W + bias ): the concat is walked as a K-split accumulate."""
        k0 = 0
        for i, a in enumerate(segs):
            last = i == len(segs) - 1
            self.gemm(a, W.row_slice(k0, a.cols), out, a.rows, out.cols, a.cols, 0, 0,
                      bias if last else None, act if last else 0, 1 if i > 0 else 0,
                      emit_split=emit_split and len(segs) == 1)
            k0 += a.cols

    # ------------------------------------------------------------------------------------------------
    # packed / length-sorted execution
    # ------------------------------------------------------------------------------------------------
    def _lens_of(self, mask):
        """Utterance lengths of a PREFIX mask (host array or device tensor), or None when it is not one."""
        if isinstance(mask, torch.Tensor) and mask.is_cuda:
            lens = getattr(mask, '_ipavsr_lens', None)      # utils/datagen.DeviceDataset.gather knows them on the host
            if lens is not None:
                return np.asarray(lens, dtype=np.int64)
            key = (mask.data_ptr(), mask._version, tuple(mask.shape))
            if key not in self._lens_cache:
                m = mask != 0
                ln = m.sum(1)
                ok = (m == (torch.arange(m.shape[1], device=m.device)[None, :] < ln[:, None])).all()
                h = torch.cat([ln.to(torch.int64), ok.to(torch.int64).reshape(1)]).cpu().numpy()   # one host sync
                if len(self._lens_cache) >= 32:
                    self._lens_cache.clear()
                self._lens_cache[key] = (h[:-1].copy() if h[-1] else None, mask)    # holds the tensor: the address stays its own
            return self._lens_cache[key][0]
        a = np.asarray(mask) != 0
        if a.ndim != 2:
            return None
        lens = a.sum(1)
        if not (a == (np.arange(a.shape[1])[None, :] < lens[:, None])).all():
            return None
        return lens.astype(np.int64)

    def _get_plan(self, inputs):
        """The _PackPlan of this batch, or None when it runs in the reference's padded layout."""
        if len(self.mask_layers) != 1 or (self.packed_mode == 'off' and not any(
                isinstance(v, Derived) for v in inputs.values())):
            return None
        ml = next(iter(self.mask_layers))
        mask = inputs.get(ml)
        if mask is None or len(mask.shape) != 2:
            return None
        N, T = int(mask.shape[0]), int(mask.shape[1])
        derived = False
        for l in self.input_layers:        # every stream must be a (N, T, F) sequence of the same batch
            if isinstance(inputs[l], Derived):
                derived = True             # computed on the device from the packed frames of its source: needs the plan
            elif l is not ml and (len(inputs[l].shape) != 3 or tuple(inputs[l].shape[:2]) != (N, T)):
                return None
        if self.packed_mode != 'force' and not derived and N * T < 2048:
            return None                    # small batches are launch-bound: the extra row movers would not pay
        lens = self._lens_of(mask)
        if lens is None or N == 0:
            return None
        if self.packed_mode != 'force' and not derived and (N * T - int(lens.sum())) * 16 < N * T:
            return None
        key = (T, lens.tobytes())
        for i, (k, plan) in enumerate(self._plans):
            if k == key:
                if i:
                    self._plans.insert(0, self._plans.pop(i))
                return plan
        plan = _PackPlan(lens, T)
        need = sum(len(a) for _, a in plan.tables)
        slot = self._pin_ring(need)
        with torch.cuda.device(self.device):
            plan.upload(self.device, slot[0])
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        slot[1] = ev
        plan.ready = ev
        self._plans.insert(0, (key, plan))
        del self._plans[8:]
        return plan

    def _pin_ring(self, need):
        """A pinned int32 staging tensor of at least `need` elements from a ring of 4: [tensor, event of the copy that last
        read it].  A slot is reused only after that copy has completed, so asynchronous H2D copies never see it change."""
        if len(self._plan_pins) < 4:
            self._plan_pins.insert(0, [torch.empty(max(need, 1 << 16), dtype=torch.int32).pin_memory(), None])
        else:
            self._plan_pins.insert(0, self._plan_pins.pop())
            if self._plan_pins[0][1] is not None:
                self._plan_pins[0][1].synchronize()
        if self._plan_pins[0][0].numel() < need:
            self._plan_pins[0][0] = torch.empty(need, dtype=torch.int32).pin_memory()
        return self._plan_pins[0]

    def _stage_targets(self, run, y):
        """Targets of a training / cost call -> device int32 BEFORE the forward kernels are enqueued, in the run's utterance
        order, through pinned staging.  (A pageable host->device copy issued after the forward pass blocks the host until
        the forward has finished on the device, and the backward is then enqueued onto an idle GPU: 1.4 ms per step.)"""
        plan = run.plan
        if y is None or isinstance(y, DevMat):
            return
        if isinstance(y, torch.Tensor) and y.is_cuda:
            if plan is None:
                run.targets_dev = y.to(torch.int32).contiguous()
            else:
                ys = y.to(torch.int32).contiguous().view(plan.N, -1)
                yd = torch.empty_like(ys)
                _lib.call('ipavsr_gather_rows', ys.data_ptr(), 4 * ys.shape[1], yd.data_ptr(), 4 * ys.shape[1],
                          4 * ys.shape[1], plan.order.data_ptr(), None, plan.N, self.stream)
                run.targets_dev = yd.view(y.shape)
            return
        yh = y.numpy() if isinstance(y, torch.Tensor) else np.asarray(y)
        if yh.dtype.kind not in 'iub':
            return                                   # regression targets (float): uploaded by the loss
        if plan is not None:
            yh = yh[plan.order_host]
        yh = np.ascontiguousarray(yh, dtype=np.int32)
        slot = self._pin_ring(yh.size)
        slot[0][:yh.size].numpy()[:] = yh.ravel()
        run.targets_dev = slot[0][:yh.size].to(self.device, non_blocking=True).view(yh.shape)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        slot[1] = ev

    def _gather(self, src_ptr, src_pitch, rows_out, cols, idx, keep=None, stream=None):
        """DevMat (rows_out x cols) = rows of a float32 matrix at `src_ptr` (row pitch src_pitch floats; device memory or
        pinned host memory) selected by the device index table `idx` (-1: zero row)."""
        out = self.new(rows_out, cols, zero=(_ld8(cols) != cols))
        _lib.call('ipavsr_gather_rows', src_ptr, 4 * src_pitch, out.ptr, 4 * out.ld, 4 * cols, idx.data_ptr(), None,
                  rows_out, stream if stream is not None else self.stream)
        return out

    def _plan_input(self, plan, l, t, stream=None):
        """Input stream `t` ((N, T, F) float32 torch tensor: device, or pinned host) in the plan's layout: packed rows for
        an encoder input, whole utterances in sorted order otherwise."""
        N, T, F = t.shape
        if l in self.pack_in and not t.is_cuda and self.host_upload == 'dma':
            # pinned host stream: one copy-engine transfer per utterance (valid frames only) — no SM is taken from the
            # compute kernels the upload overlaps with
            out = self.new(plan.M + 1, F, zero=(_ld8(F) != F))
            _lib.call('ipavsr_upload_ragged', t.data_ptr(), 4 * T * F, 4 * F, out.ptr, 4 * out.ld,
                      plan.order32.ctypes.data_as(C.c_void_p), plan.offsets_host.ctypes.data_as(C.c_void_p), N,
                      stream if stream is not None else self.stream)
            return out
        if l in self.pack_in:
            return self._gather(t.data_ptr(), F, plan.M + 1, F, plan.pack, stream=stream)
        return self._gather(t.data_ptr(), F, N * T, F, plan.perm, stream=stream)

    @staticmethod
    def _as_f32_tensor(arr):
        """(tensor, pinned-or-device flag) of an input array without copying when it already is float32 contiguous."""
        if isinstance(arr, torch.Tensor):
            t = arr
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(arr).astype(np.float32, copy=False)))
        if t.dtype != torch.float32 or not t.is_contiguous():
            t = t.to(torch.float32).contiguous()
        return t, (t.is_cuda or t.is_pinned())

    # ------------------------------------------------------------------------------------------------
    # input staging
    # ------------------------------------------------------------------------------------------------
    def _upload(self, arr, kind):
        if isinstance(arr, torch.Tensor):
            t = arr
        else:
            a = np.asarray(arr)
            if kind == 'mask':
                a = np.ascontiguousarray(a.astype(np.uint8, copy=False))
            elif kind == 'int':
                a = np.ascontiguousarray(a.astype(np.int32, copy=False))
            else:
                a = np.ascontiguousarray(a.astype(np.float32, copy=False))       # allow_input_downcast
            t = torch.from_numpy(a)
        if kind == 'mask':
            return t.to(torch.uint8).to(self.device, non_blocking=True).contiguous()
        if kind == 'int':
            return t.to(torch.int32).to(self.device, non_blocking=True).contiguous()
        t = t.to(torch.float32)
        N, T, F = t.shape
        ld = _ld8(F)
        if ld == F:
            d = t.to(self.device, non_blocking=True).contiguous()
            return DevMat(d, d.data_ptr(), N * T, F, F)
        d = torch.zeros(N * T, ld, dtype=torch.float32, device=self.device)
        d[:, :F].copy_(t.reshape(N * T, F), non_blocking=True)
        return DevMat(d, d.data_ptr(), N * T, F, ld)

    @staticmethod
    def _host_signature(inputs, layers):
        sig = []
        for l in layers:
            a = inputs[l]
            if isinstance(a, Derived):
                sig.append((id(l), id(a.source), type(a).__name__))
                continue
            if isinstance(a, torch.Tensor):
                if a.is_cuda:
                    continue
                sig.append((id(l), a.data_ptr(), tuple(a.shape)))
            else:
                a = np.asarray(a)
                sig.append((id(l), a.__array_interface__['data'][0], tuple(a.shape)))
        return tuple(sig)

    def prefetch(self, inputs):
        """Stage the host inputs of a later forward() now (function(...).prefetch)."""
        plan = self._get_plan(inputs)
        staged = self._stage_inputs(inputs, allow_chunks=False, plan=plan)
        if staged:
            self._prefetched.append((self._host_signature(inputs, self.input_layers), (staged, plan)))
            del self._prefetched[:-4]          # a forgotten prefetch must not pin device memory for ever

    def _take_prefetched(self, inputs):
        if not self._prefetched:
            return None
        sig = self._host_signature(inputs, self.input_layers)
        for i, (s, staged) in enumerate(self._prefetched):
            if s == sig:
                del self._prefetched[i]
                return staged
        return None

    def _upload_chunked(self, arr, cs, nchunks=4):
        """(N,T,F) host array -> device matrix filled by `nchunks` asynchronous row-chunk copies on stream `cs`, each with
        its own event (DevMat.chunks).  None when the input is too small or needs a padded leading dimension."""
        t = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(arr).astype(np.float32, copy=False)))
        if t.dtype != torch.float32 or t.dim() != 3 or not t.is_contiguous():
            return None
        N, T, F = t.shape
        if F % 8 != 0 or N * T < 8192 or N < 2 * nchunks:
            return None
        d = torch.empty(N * T, F, dtype=torch.float32, device=self.device)
        h2 = t.reshape(N * T, F)
        chunks, per = [], (N + nchunks - 1) // nchunks
        for c in range(nchunks):
            r0, r1 = c * per * T, min(N, (c + 1) * per) * T
            if r1 <= r0:
                break
            d[r0:r1].copy_(h2[r0:r1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(cs)
            chunks.append((r0, r1 - r0, ev))
        return DevMat(d, d.data_ptr(), N * T, F, F, chunks=chunks)

    def _stage_inputs(self, inputs, allow_chunks=True, plan=None):
        """Host inputs are copied on a dedicated copy stream, all issued up front in graph order, so that the upload of
        the later streams overlaps the encoder of the first ones; the compute stream waits per input, on first use.
        Pinned host tensors make the copies truly asynchronous.  Returns {layer: (device value, event)}."""
        host = [l for l in self.input_layers
                if not (isinstance(inputs[l], torch.Tensor) and inputs[l].is_cuda) and not isinstance(inputs[l], Derived)]
        if not host or os.environ.get('IPAVSR_COPY_STREAM', '1') == '0':
            return {}
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        cs = self._copy_stream
        # The staging buffers are allocated with the copy stream current: they come from that stream's pool of the caching
        # allocator, which hands a block out again only after every stream it was recorded on has passed its last use —
        # the copies need not wait for the compute stream (they did, which kept an upload issued behind a step's kernels
        # from overlapping them).  What the copies do read from the compute stream is the plan's index tables.
        if plan is not None and plan.ready is not None:
            cs.wait_event(plan.ready)
        staged = {}
        # the mask first (tiny, needed by every LSTM), then the streams in graph order
        order = [l for l in host if l in self.mask_layers] + [l for l in host if l not in self.mask_layers]
        with torch.cuda.stream(cs):
            for l in order:
                val = None
                if plan is not None:
                    if l in self.mask_layers:
                        continue                    # the plan carries the (sorted) mask
                    # pinned host memory is read by the gather kernel itself: only the valid frames of an encoder
                    # stream cross PCIe (ragged upload); pageable memory is copied whole first
                    t, direct = self._as_f32_tensor(inputs[l])
                    if not direct:
                        t = t.to(self.device, non_blocking=True)
                    val = self._plan_input(plan, l, t, stream=C.c_void_p(cs.cuda_stream))
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    val.t.record_stream(main)
                    staged[l] = (val, ev, t)        # t: keeps the source alive until the call consumed the staged value
                    continue
                if allow_chunks and l in self.chunkable and self.gemm_mode == 4 and l not in self.mask_layers:
                    val = self._upload_chunked(inputs[l], cs)
                if val is None:
                    val = self._upload(inputs[l], 'mask' if l in self.mask_layers else 'float')
                ev = torch.cuda.Event()
                ev.record(cs)
                (val if isinstance(val, torch.Tensor) else val.t).record_stream(main)
                staged[l] = (val, ev)
        return staged

    def _input(self, run, l, inputs, staged):
        """Value of InputLayer `l` for this run (memoised in run.vals): staged host copy, device tensor, the plan's mask, or a
        stream derived on the device from another input (derived.py)."""
        plan = run.plan
        N, T = run.N, run.T
        a = inputs[l]
        if isinstance(a, Derived):
            if plan is None:
                raise ValueError('a derived input stream needs the utterance lengths: pass a prefix mask (one mask input)')
            run.vals[l] = [self._derive(run, l, a, inputs, staged)]
        elif plan is not None and l in self.mask_layers:
            run.vals[l] = plan.mask                 # the mask in sorted utterance order (part of the plan)
        elif l in staged:
            val, ev = staged[l][:2]
            if not (isinstance(val, DevMat) and val.chunks):     # chunked inputs are awaited chunk by chunk
                torch.cuda.current_stream(self.device).wait_event(ev)
            run.vals[l] = val if l in self.mask_layers else [val]
            if plan is not None and l in self.pack_in:
                run.packed.add(l)
        elif plan is not None:
            t, direct = self._as_f32_tensor(a)
            if not direct:
                t = t.to(self.device, non_blocking=True)
            run.vals[l] = [self._plan_input(plan, l, t)]
            run.keep.append(t)
            if l in self.pack_in:
                run.packed.add(l)
        elif l in self.mask_layers:
            run.vals[l] = self._upload(a, 'mask')
        else:
            run.vals[l] = [self._upload(a, 'float')]
        return run.vals[l]

    def _packed_rows(self, run, l, inputs, staged):
        """The M valid frames (+ the zero row) of input stream `l` in the plan's sorted order."""
        plan = run.plan
        if l not in run.vals:
            self._input(run, l, inputs, staged)
        v = run.vals[l][0]
        if l in run.packed:
            return v
        key = ('packed', l)
        if key not in run.saved:
            out = self.new(plan.M + 1, v.cols, zero=(_ld8(v.cols) != v.cols))
            _lib.call('ipavsr_gather_rows', v.ptr, 4 * v.ld, out.ptr, 4 * out.ld, 4 * v.cols, plan.valid.data_ptr(), None,
                      plan.M, self.stream)
            _lib.call('ipavsr_fill', out.ptr + 4 * plan.M * out.ld, out.ld, 0.0, self.stream)     # row M: the zero row
            run.saved[key] = out
        return run.saved[key]

    def _derive(self, run, l, spec, inputs, staged):
        """Computes a derived stream (derived.py) on the packed frames of its source and returns it in the layout `l` takes:
        packed rows for an encoder input, the padded sorted layout (zero padding) otherwise."""
        plan, st = run.plan, self.stream
        src_layer = None
        for cand in self.input_layers:
            if cand is not l and inputs[cand] is spec.source:
                src_layer = cand
        if src_layer is None or isinstance(inputs[src_layer], Derived):
            raise ValueError('the source of a derived stream must be the array passed for another input of the same call')
        x = self._packed_rows(run, src_layer, inputs, staged)
        M, N = plan.M, plan.N
        if isinstance(spec, DiffImages):
            y = self.new(M + 1, x.cols, zero=(_ld8(x.cols) != x.cols))
            _lib.call('ipavsr_diff_image', x.ptr, x.ld, y.ptr, y.ld, plan.offsets.data_ptr(), N, x.cols, st)
        elif isinstance(spec, DctFeatures):
            D, K = x.cols, spec.no_coeff
            if spec.image_shape[0] * spec.image_shape[1] != D:
                raise ValueError('cannot reshape frames of %d pixels into %r' % (D, spec.image_shape))
            key = (spec.image_shape, K)
            if key not in self._dct_basis:
                from .utils.preprocessing import zigzag_order
                cols = torch.from_numpy(np.ascontiguousarray(zigzag_order(*spec.image_shape)[1:K + 1],
                                                             dtype=np.int32)).to(self.device)
                ldb = (K + 3) // 4 * 4
                basis = torch.empty(D, ldb, dtype=torch.float32, device=self.device)
                _lib.call('ipavsr_dct_basis', basis.data_ptr(), ldb, cols.data_ptr(), D, K, st)
                self._dct_basis[key] = (basis, ldb)
            basis, ldb = self._dct_basis[key]
            c = self.new(M + 1, K, zero=(_ld8(K) != K))
            for f0 in range(0, M, 65535 * 128):
                n = min(65535 * 128, M - f0)
                _lib.call('ipavsr_dct_project', x.ptr + 4 * x.ld * f0, x.ld, basis.data_ptr(), ldb, c.ptr + 4 * c.ld * f0,
                          c.ld, n, D, K, st)
            if spec.deltas:
                y = self.new(M + 1, 3 * K, zero=(_ld8(3 * K) != 3 * K))
                _lib.call('ipavsr_deltas_fir_f32', c.ptr, c.ld, y.ptr, y.ld, plan.offsets.data_ptr(), N, K, spec.window,
                          max(int(plan.lens_sorted.max()), 1), st)
            else:
                y = c
        else:
            raise TypeError('unknown derived stream %r' % (type(spec).__name__,))
        _lib.call('ipavsr_fill', y.ptr + 4 * M * y.ld, y.ld, 0.0, st)        # the zero row (no utterance owns it)
        if l in self.pack_in:
            run.packed.add(l)
            return y
        full = self.new(N * run.T, y.cols, zero=(_ld8(y.cols) != y.cols))
        _lib.call('ipavsr_gather_rows', y.ptr, 4 * y.ld, full.ptr, 4 * full.ld, 4 * y.cols, plan.unpack.data_ptr(), None,
                  N * run.T, st)
        return full

    # ------------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------------
    def forward(self, inputs, window, deterministic=True, train=False, dropout_masks=None, update_bn=True, targets=None):
        """inputs: {InputLayer: array}.  Returns (_Run, output Val).  `train` keeps what backward needs."""
        lib, st = self.lib, self.stream
        first = None
        for l in self.input_layers:
            if l not in self.mask_layers and not isinstance(inputs[l], Derived):
                first = inputs[l]
                break
        N, T = int(first.shape[0]), int(first.shape[1])
        run = _Run(N, T)
        self._zpool = self._opool = None    # small zeroed / one-initialised scratch: fresh pools per step
        self.arena.sync_from_peers()
        pre = self._take_prefetched(inputs)
        if pre is not None:
            staged, plan = pre
        else:
            plan = self._get_plan(inputs)
            staged = self._stage_inputs(inputs, plan=plan)
        run.plan = plan
        run.targets_dev = None
        self._stage_targets(run, targets)
        self._split_cache = {}
        self._split_by_storage = {}
        self._amax = {}
        if self.gemm_mode in (1, 4):
            self._refresh_param_split()
        run.window = int(window) if window is not None else 0
        run.deterministic = deterministic
        ar = self.arena
        self._pin_stream(True)
        run.main_stream = torch.cuda.current_stream(self.device)
        run.use_branches = self._use_branches(N, T)
        if run.use_branches:
            # everything the branches share is created on the trunk's stream BEFORE they fork from it: the scratch pools,
            # the [1.0, 14] constant, the buffers several LSTMs write their column slice of
            self._zpool, self._zpos = torch.zeros(16384, dtype=torch.float32, device=self.device), 0
            self._opool = None
            self._ones2()
            self._const14()          # (made once, outside any graph capture: it takes a host scalar)
            for cl, (_, total) in self.cat_plan.items():
                run.cat[cl] = (self.new(N * T, total, zero=(_ld8(total) != total)), self._ones2())
        try:
            for l in self.layers:
                if _Nvtx.on:
                    torch.cuda.nvtx.range_push('fwd %s' % (l.name or type(l).__name__))
                b = self._branch_of.get(l) if run.use_branches else None
                if b is not None:
                    self._branch_enter(run, b, False)
                    try:
                        self._forward_layer(run, l, inputs, staged, plan, deterministic, train, dropout_masks, update_bn)
                    finally:
                        self._branch_leave(run)
                    run.branch_done[b] = run.branch_done.get(b, 0) + 1
                else:
                    if run.use_branches:
                        # a trunk layer waits for the branches it reads from (once per batch of their layers)
                        for i in (getattr(l, 'input_layers', None) or [getattr(l, 'input_layer', None)]):
                            bi = self._branch_of.get(i) if i is not None else None
                            if bi is not None and run.branch_joined.get(bi, 0) < run.branch_done.get(bi, 0):
                                self._branch_join(run, bi)
                                run.branch_joined[bi] = run.branch_done[bi]
                    self._forward_layer(run, l, inputs, staged, plan, deterministic, train, dropout_masks, update_bn)
                if _Nvtx.on:
                    torch.cuda.nvtx.range_pop()
            if run.use_branches:
                for b in sorted(run.branches_used):       # outputs taken from inside a branch; nothing left running
                    self._branch_join(run, b)
            return self._forward_finish(run)
        finally:
            if run.use_branches:
                self._branch_leave(run)
            self._pin_stream(False)

    def _forward_layer(self, run, l, inputs, staged, plan, deterministic, train, dropout_masks, update_bn):
        lib, st, ar = self.lib, self.stream, self.arena
        N, T = run.N, run.T
        if True:
            for i in (getattr(l, 'input_layers', None) or [getattr(l, 'input_layer', None)]):
                if i is not None and i in run.pending and not (isinstance(l, L.LSTMLayer) and i in self.mask_layers):
                    self._wait(run, i)
            if isinstance(l, L.InputLayer):
                if l not in run.vals:
                    self._input(run, l, inputs, staged)
            elif isinstance(l, L.ReshapeLayer):
                run.vals[l] = run.vals[l.input_layer]
                if l.input_layer in run.packed:
                    run.packed.add(l)
            elif isinstance(l, L.DenseLayer):
                segs = run.vals[l.input_layer]
                rows = segs[0].rows
                if l.input_layer in run.packed and l not in self.pack_tail:
                    run.packed.add(l)
                W = ar.mat((l, 'W'))
                b = ar.mat((l, 'b')).ptr if l.b is not None else None
                out = self.new(rows, l.num_units)
                if l.nonlinearity.name == 'softmax':
                    logits = self.new(rows, l.num_units)
                    self._proj(segs, W, logits, b, 0)
                    _lib.call('ipavsr_softmax', logits.ptr, logits.ld, out.ptr, out.ld, rows, l.num_units, st)
                elif len(segs) == 1 and segs[0].chunks and self.gemm_mode == 4:
                    # the input is still arriving: one product per row chunk, each split with its own scale
                    x = segs[0]
                    amax_t = self._zeros2()
                    for (r0, n, ev) in x.chunks:
                        torch.cuda.current_stream(self.device).wait_event(ev)
                        self.gemm(x.row_slice(r0, n), W, out.row_slice(r0, n), n, out.cols, x.cols, 0, 0, b,
                                  ACT[l.nonlinearity.name], 0, amax_t=amax_t)
                    self._set_amax(out, amax_t)
                else:
                    self._proj(segs, W, out, b, ACT[l.nonlinearity.name], emit_split=True)
                if l.input_layer in run.packed and l in self.pack_tail:
                    # last layer of a packed encoder: expand to the padded (sorted) layout; every padding frame takes the
                    # output of the zero row, enc(0).  Backward needs the packed output.
                    run.saved[l] = out
                    full = self.new(N * T, l.num_units, zero=(out.ld != l.num_units))
                    _lib.call('ipavsr_gather_rows', out.ptr, 4 * out.ld, full.ptr, 4 * full.ld, 4 * l.num_units,
                              plan.unpack.data_ptr(), None, N * T, st)
                    out = full
                run.vals[l] = [out]
            elif isinstance(l, L.BatchNormLayer):
                x = run.vals[l.input_layer][0]
                F = x.cols
                y = self.new(x.rows, F)
                beta, gamma = ar.mat((l, 'beta')).ptr, ar.mat((l, 'gamma')).ptr
                rm, ri = ar.aux_ptr((l, 'mean')), ar.aux_ptr((l, 'inv_std'))
                if deterministic:
                    _lib.call('ipavsr_bn_fwd', x.ptr, x.ld, y.ptr, y.ld, beta, gamma, rm, ri, None, None, None,
                              x.rows, F, 0, l.epsilon, l.alpha, 1, 0, st)
                else:
                    stats = torch.empty(2 * F, dtype=torch.float64, device=self.device)
                    save = torch.empty(2 * F, dtype=torch.float32, device=self.device)
                    _lib.call('ipavsr_bn_stats', x.ptr, x.ld, stats.data_ptr(), x.rows, F, st)
                    m_total = x.rows
                    if self.world is not None:
                        torch.distributed.all_reduce(stats, group=self.world[2])
                        m_total = x.rows * self.world[1]
                    _lib.call('ipavsr_bn_fwd', x.ptr, x.ld, y.ptr, y.ld, beta, gamma, rm, ri, stats.data_ptr(),
                              save.data_ptr(), save.data_ptr() + 4 * F, x.rows, F, m_total, l.epsilon, l.alpha, 0,
                              1 if update_bn else 0, st)
                    run.saved[l] = (save, m_total)
                    if update_bn:
                        ar.touch([l.mean, l.inv_std])
                run.vals[l] = [y]
            elif isinstance(l, L.DropoutLayer):
                segs = run.vals[l.input_layer]
                if deterministic:
                    run.vals[l] = segs
                else:
                    rows = segs[0].rows
                    tot = sum(s.cols for s in segs)
                    if dropout_masks is not None and l.name in dropout_masks:
                        km = np.asarray(dropout_masks[l.name]).reshape(rows, tot)
                        if plan is not None:
                            km = km[plan.perm_host if rows == N * T else plan.order_host]
                        keep_full = torch.from_numpy(np.ascontiguousarray(km.astype(np.uint8))).to(self.device)
                    else:
                        keep_full = torch.empty(rows, tot, dtype=torch.uint8, device=self.device)
                        # data-parallel shards draw from different streams (rank mixed into the seed): the masks of the
                        # global batch are then independent, as in the single-process step
                        _lib.call('ipavsr_dropout_mask', keep_full.data_ptr(), rows * tot, float(l.p),
                                  self.dropout_seed + 7919 * (self.world[0] if self.world is not None else 0),
                                  self.dropout_calls << 32, st)
                        self.dropout_calls += 1
                    scale = 1.0 / (1.0 - l.p) if l.rescale else 1.0
                    outs, keeps, c0 = [], [], 0
                    for s in segs:
                        k = keep_full[:, c0:c0 + s.cols].contiguous()
                        o = self.new(rows, s.cols)
                        _lib.call('ipavsr_dropout', s.ptr, s.ld, k.data_ptr(), o.ptr, o.ld, rows, s.cols, scale, st)
                        outs.append(o)
                        keeps.append(k)
                        c0 += s.cols
                    run.saved[l] = (keeps, scale)
                    run.vals[l] = outs
            elif isinstance(l, L.DeltaLayer):
                x = self._single(run.vals[l.input_layer])
                y = self.new(x.rows, 3 * x.cols)
                _lib.call('ipavsr_delta_fwd', x.ptr, x.ld, y.ptr, y.ld, N, T, x.cols, run.window, self.delta_exact, st)
                run.vals[l] = [y]
            elif isinstance(l, L.LSTMLayer):
                segs = run.vals[l.input_layers[0]]
                H = l.num_units
                if l.mask_incoming_index > 0:
                    mask = run.vals[l.input_layers[1]]
                else:
                    mask = torch.ones(N, T, dtype=torch.uint8, device=self.device)
                # LSTMs that read the same input (the two directions of the aggregate BLSTM) get ALL their input projections
                # before the first recurrence is launched: a recurrence holds 128 SMs for its whole latency chain, and a
                # projection GEMM issued behind it crawls on the 20 SMs left (241 us instead of 90 in the 512-utterance
                # step) and delays its own recurrence by that much
                # (their output / state buffers too: the zero fills of the padding columns would wait behind the recurrence)
                def prepare(sl):
                    Hs = sl.num_units
                    xs = self.new(N * T, 4 * Hs)
                    self._proj(segs, ar.mat((sl, 'W_in')), xs, ar.mat((sl, 'b')).ptr, 0)
                    cb = None
                    if sl in self.cat_of:
                        cl, coff = self.cat_of[sl]
                        if cl not in run.cat:
                            total = self.cat_plan[cl][1]
                            bound = self._ones2()
                            run.cat[cl] = (self.new(N * T, total, zero=(_ld8(total) != total)), bound)
                        cat, cb = run.cat[cl]
                        o = DevMat(cat.t, cat.ptr + 4 * coff, N * T, Hs, cat.ld)
                    else:
                        o = self.new(N * T, Hs, zero=(_ld8(Hs) != Hs))
                    g = c = hp = None
                    if train:
                        g = self.new(N * T, 4 * Hs)
                        c = self.new(N * T, Hs)
                        if o.ld != _ld8(Hs):
                            # the kernels share one leading dimension between out and hprev: give hprev the concat's
                            t_hp = torch.empty(N * T * o.ld, dtype=torch.float32, device=self.device)
                            hp = DevMat(t_hp, t_hp.data_ptr(), N * T, Hs, o.ld)
                        else:
                            hp = self.new(N * T, Hs, zero=(_ld8(Hs) != Hs))
                        if c.ld != Hs:       # cell is dense (ld = H) inside the kernels
                            c = DevMat(c.t, c.ptr, N * T, Hs, Hs)
                    return xs, cb, o, g, c, hp

                if l not in run.xw_ready:
                    for sl in (l,) + self._lstm_siblings.get(l, ()):
                        if sl not in run.xw_ready:
                            run.xw_ready[sl] = prepare(sl)
                xw, cat_bound, out, gates, cell, hprev = run.xw_ready.pop(l)
                peep = ar.mat((l, 'peep')).ptr if l.peepholes else None
                nbytes = lib.ipavsr_lstm_workspace_bytes(N, T, H)
                if self.concurrent_lstm:
                    ws = torch.empty((int(nbytes) + 3) // 4, dtype=torch.float32, device=self.device) \
                        if self.lstm_impl != 0 else self._workspace(16)
                    run.keep.append(ws)
                    side, sh = self._fork()
                else:
                    ws, side, sh = self._workspace(nbytes), None, st
                whid = ar.mat((l, 'W_hid'))
                if (self.gemm_mode == 4 and self.lstm_impl == 0 and
                        lib.ipavsr_lstm_fwd_f16_supported(N, T, H, whid.ld)):
                    # tensor-core recurrence on the fp16 split of W_hid that the parameter arena already carries
                    wh, wl, we = self._split16(whid)
                    _lib.call('ipavsr_lstm_fwd_f16', xw.ptr, wh, wl, we, whid.ld, peep, ar.mat((l, 'cell_init')).ptr,
                              ar.mat((l, 'hid_init')).ptr, mask.data_ptr(), out.ptr,
                              gates.ptr if gates else None, cell.ptr if cell else None, hprev.ptr if hprev else None,
                              N, T, H, out.ld, 1 if l.backwards else 0, sh)
                elif (self.gemm_mode == 4 and self.lstm_impl == 0 and H > 256 and
                      lib.ipavsr_lstm_steps_supported(N, T, H, whid.ld)):
                    # wide layers (H = 500): one tensor-core GEMM + one cell kernel per time step (csrc/lstm_steps_tc.cu)
                    wh, wl, we = self._split16(whid)
                    sbytes = int(lib.ipavsr_lstm_steps_workspace_bytes(N, T, H))
                    sws = torch.empty((sbytes + 3) // 4, dtype=torch.float32, device=self.device)
                    run.keep.append(sws)
                    act = plan.active_rows.ctypes.data_as(C.c_void_p) if plan is not None else None
                    _lib.call('ipavsr_lstm_fwd_f16_steps', xw.ptr, whid.ptr, wh, wl, we, whid.ld, peep,
                              ar.mat((l, 'cell_init')).ptr, ar.mat((l, 'hid_init')).ptr, mask.data_ptr(), out.ptr,
                              gates.ptr if gates else None, cell.ptr if cell else None, hprev.ptr if hprev else None,
                              N, T, H, out.ld, 1 if l.backwards else 0, act, sws.data_ptr(), sbytes, sh)
                else:
                    _lib.call('ipavsr_lstm_fwd', xw.ptr, whid.ptr, peep, ar.mat((l, 'cell_init')).ptr,
                              ar.mat((l, 'hid_init')).ptr, mask.data_ptr(), out.ptr,
                              gates.ptr if gates else None, cell.ptr if cell else None, hprev.ptr if hprev else None,
                              N, T, H, out.ld, 1 if l.backwards else 0, self.lstm_impl, ws.data_ptr(),
                              int(nbytes) if (self.lstm_impl != 0 or not self.concurrent_lstm) else 16, sh)
                if side is not None:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    run.pending[l] = ev
                run.keep.append(xw)
                run.saved[l] = (mask, gates, cell, hprev)
                run.vals[l] = [out]
                if self.gemm_mode == 4:
                    # h = o * tanh(c) lies in (-1, 1); padded / first steps carry hid_init: |out|, |hprev| <= max(1, |hid_init|)
                    # is known without a pass over the (N*T, H) tensors
                    bound = self._ones2()
                    _lib.call('ipavsr_amax', ar.mat((l, 'hid_init')).ptr, H, 1, H, bound.data_ptr(), st)
                    if cat_bound is not None:
                        # the materialised concat is bounded by the largest of its LSTMs' bounds (atomic max, same stream)
                        _lib.call('ipavsr_amax', ar.mat((l, 'hid_init')).ptr, H, 1, H, cat_bound.data_ptr(), st)
                    else:
                        self._set_amax(out, bound)
                    if hprev is not None:
                        self._set_amax(hprev, bound)       # same bound, same scale exponent: the two splits share the pair
            elif isinstance(l, (L.ElemwiseSumLayer, L.AdaptiveElemwiseSumLayer)):
                ins = [self._single(run.vals[i]) for i in l.input_layers]
                rows, F = ins[0].rows, ins[0].cols
                out = self.new(rows, F, zero=(_ld8(F) != F))
                ptrs = (C.c_void_p * len(ins))(*[i.ptr for i in ins])
                lds = (C.c_int * len(ins))(*[i.ld for i in ins])
                coeffs = ar.mat((l, 'coeffs')).ptr if isinstance(l, L.AdaptiveElemwiseSumLayer) else None
                _lib.call('ipavsr_fuse_sum', ptrs, lds, len(ins), coeffs, out.ptr, out.ld, rows, F, st)
                run.vals[l] = [out]
            elif isinstance(l, L.ConcatLayer):
                if l in run.cat:
                    cat, bound = run.cat[l]
                    if self.gemm_mode == 4:
                        self._set_amax(cat, bound)
                    run.vals[l] = [cat]
                else:
                    segs = []
                    for i in l.input_layers:
                        segs.extend(run.vals[i])
                    run.vals[l] = segs
            elif isinstance(l, L.SliceLayer):
                x = self._single(run.vals[l.input_layer])
                out = self.new(N, x.cols, zero=(_ld8(x.cols) != x.cols))
                _lib.call('ipavsr_slice_last', x.ptr, x.ld, out.ptr, out.ld, N, T, x.cols, 0, 0, st)
                run.vals[l] = [out]
            else:
                raise TypeError('unsupported layer type %s' % type(l).__name__)

    def _forward_finish(self, run):
        plan, st = run.plan, self.stream
        N, T = run.N, run.T
        self._join(run)
        outv = run.vals[self.out]
        run.out_sorted = outv
        if plan is not None and self.out not in run.packed:
            # back to the caller's utterance order
            res = []
            for sg in outv:
                if sg.rows == N * T:
                    idx = plan.unperm
                elif sg.rows == N:
                    idx = plan.inv
                else:
                    raise RuntimeError('output with %d rows in a batch of %d x %d' % (sg.rows, N, T))
                o = self.new(sg.rows, sg.cols, zero=(_ld8(sg.cols) != sg.cols))
                _lib.call('ipavsr_gather_rows', sg.ptr, 4 * sg.ld, o.ptr, 4 * o.ld, 4 * sg.cols, idx.data_ptr(), None,
                          sg.rows, st)
                res.append(o)
            outv = res
        return run, outv

    def _single(self, segs):
        if len(segs) == 1:
            return segs[0]
        rows = segs[0].rows
        tot = sum(s.cols for s in segs)
        out = self.new(rows, tot, zero=True)
        c0 = 0
        for s in segs:
            _lib.call('ipavsr_copy2d', s.ptr, s.ld, out.ptr + 4 * c0, out.ld, rows, s.cols, None, 0, self.stream)
            c0 += s.cols
        return out

    def read_device(self, segs):
        """Val -> contiguous (rows, cols) float32 torch tensor on the device."""
        parts = [s.torch_view() for s in segs]
        return (parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)).contiguous()

    def read(self, segs):
        """Val -> host ndarray (rows, cols)."""
        return np.concatenate([s.torch_view().detach().cpu().numpy() for s in segs], axis=1)

    # ------------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------------
    def _grad_target(self, run, layer, like):
        """Buffers to write d(layer output) into: (segs, accumulate flag)."""
        if layer in run.grads:
            segs, owned = run.grads[layer]
            if not owned:
                new = [self.new(s.rows, s.cols, zero=(_ld8(s.cols) != s.cols)) for s in segs]
                for a, b in zip(segs, new):
                    _lib.call('ipavsr_copy2d', a.ptr, a.ld, b.ptr, b.ld, a.rows, a.cols, None, 0, self.stream)
                run.grads[layer] = (new, True)
                segs = new
            return segs, 1
        segs = [self.new(s.rows, s.cols, zero=(_ld8(s.cols) != s.cols)) for s in like]
        run.grads[layer] = (segs, True)
        return segs, 0

    def _pass_grad(self, run, layer, segs):
        if layer is None or not self.requires_grad.get(layer, False):
            return
        if layer not in run.grads:
            run.grads[layer] = (segs, False)
            return
        dst, _ = self._grad_target(run, layer, segs)
        for a, b in zip(segs, dst):
            _lib.call('ipavsr_copy2d', a.ptr, a.ld, b.ptr, b.ld, a.rows, a.cols, None, 1, self.stream)

    def backward(self, run, dlogits, softmax_head=True):
        """dlogits: DevMat gradient w.r.t. the pre-softmax logits of the output Dense (softmax_head) or w.r.t. the output of
        a non-softmax output Dense (regression nets: auto-encoder fine-tuning).  Writes the gradient arena."""
        lib, st, ar = self.lib, self.stream, self.arena
        N, T = run.N, run.T
        G = lambda key: ar.mat(key, 'grad')
        head = self.out
        while isinstance(head, L.ReshapeLayer):
            head = head.input_layer
        if softmax_head and not (isinstance(head, L.DenseLayer) and head.nonlinearity.name == 'softmax'):
            raise ValueError('the network output must be a softmax DenseLayer to train')
        if not softmax_head and not (isinstance(head, L.DenseLayer) and head.nonlinearity.name != 'softmax'):
            raise ValueError('a squared-error objective needs a non-softmax DenseLayer output')
        run.grads[head] = ([dlogits], True)
        self._cur_run = run
        self._ar_works, self._ar_hi = [], (ar.flat.numel() if (self.world is not None and self.ar_overlap and
                                                                getattr(self, '_ar_enabled', True)) else None)
        # number of not-yet-processed consumers of every layer: when it reaches zero the layer's gradient is final and
        # an LSTM recurrence can be launched ahead of the walk on a side stream
        remaining = {}
        for l in self.layers:
            for i in (getattr(l, 'input_layers', None) or [getattr(l, 'input_layer', None)]):
                if i is not None:
                    remaining[i] = remaining.get(i, 0) + 1
        for l in self._bwd_order:
            in_layers = [i for i in (getattr(l, 'input_layers', None) or [getattr(l, 'input_layer', None)])
                         if i is not None]
            if run.use_branches:
                # small batch: the layers of an input branch run their backward on the branch's stream (ordered behind the
                # trunk where the gradient comes from it); the trunk's own layers on the trunk's
                self._branch_leave(run)
                b = self._branch_of.get(l)
                if b is not None and l in run.grads and not isinstance(l, L.InputLayer):
                    self._branch_enter(run, b, l in self._trunk_fed)
                st = self.stream
            if l not in run.grads or isinstance(l, L.InputLayer):
                self._release(run, in_layers, remaining)
                continue
            if _Nvtx.on:
                torch.cuda.nvtx.mark('bwd %s' % (l.name or type(l).__name__))
            gsegs, _ = run.grads[l]
            if isinstance(l, L.ReshapeLayer):
                self._pass_grad(run, l.input_layer, gsegs)
            elif isinstance(l, L.DenseLayer):
                dY = gsegs[0]
                xin = run.vals[l.input_layer]
                rows, Nout = dY.rows, l.num_units
                if l.nonlinearity.name == 'softmax':
                    if l is not head:
                        raise ValueError('softmax DenseLayer %r is not the network head: its gradient is only defined '
                                         'through the fused loss kernels' % (l.name,))
                    dZ = dY          # the loss kernel already went through the softmax
                    if l.b is not None:
                        _lib.call('ipavsr_colsum', dZ.ptr, dZ.ld, G((l, 'b')).ptr, rows, Nout, 0, st)
                else:
                    owned = run.grads[l][1]
                    y = run.vals[l][0]
                    if l in self.pack_tail and l.input_layer in run.packed:
                        # backward of the expansion: valid frames gather their gradient rows, the zero row collects the
                        # gradients of all padding frames (they share its output)
                        plan = run.plan
                        y = run.saved[l]
                        dYp = self.new(plan.M + 1, Nout, zero=(_ld8(Nout) != Nout))
                        _lib.call('ipavsr_gather_rows', dY.ptr, 4 * dY.ld, dYp.ptr, 4 * dYp.ld, 4 * Nout,
                                  plan.valid.data_ptr(), None, plan.M, st)
                        _lib.call('ipavsr_colsum_masked', dY.ptr, dY.ld, plan.mask.data_ptr(), 1,
                                  dYp.ptr + 4 * plan.M * dYp.ld, N * T, Nout, 0, st)
                        dY, owned, rows = dYp, True, plan.M + 1
                    need_dx = self.requires_grad.get(l.input_layer, False)
                    fused = (self.gemm_mode == 4 and self.fuse_prep and not any(a.chunks for a in xin) and
                             lib.ipavsr_dense_bwd_prep_f16_supported(dY.ptr, dY.ld, y.ptr, y.ld, Nout, 16, 16, dY.ld) and
                             all(lib.ipavsr_gemm_f16_supported(a.cols, Nout, rows, 16, a.ld, 16, dY.ld) and
                                 (not need_dx or lib.ipavsr_gemm_f16_supported(rows, a.cols, Nout, 16, dY.ld, 16,
                                                                               ar.mat((l, 'W')).ld)) for a in xin))
                    if fused:
                        # f16x3 mode: dZ = dY act'(Y) leaves the prep kernel only as the fp16 hi/lo operand pair of the two
                        # GEMMs below (+ the bias gradient); its scale comes from the |dY|max the producing dgrad epilogue
                        # left behind (|act'| <= 1), else from one max pass over dY
                        bnd = self._amax.pop((dY.ptr, dY.rows, dY.cols, dY.ld), None)
                        if bnd is None:
                            bt = self._zeros2()
                            _lib.call('ipavsr_amax', dY.ptr, dY.ld, rows, Nout, bt.data_ptr(), st)
                        else:
                            bt = bnd[0]
                        n16 = max(rows * dY.ld, 8)
                        zhi = torch.empty(n16, dtype=torch.float16, device=self.device)
                        zlo = torch.empty(n16, dtype=torch.float16, device=self.device)
                        zex = self._zeros2()
                        _lib.call('ipavsr_dense_bwd_prep_f16', dY.ptr, dY.ld, y.ptr, y.ld,
                                  G((l, 'b')).ptr if l.b is not None else None, rows, Nout, ACT[l.nonlinearity.name], 0,
                                  bt.data_ptr(), zhi.data_ptr(), zlo.data_ptr(), dY.ld, zex.data_ptr() + 4, st)
                        # a handle for dZ that only exists as its split: the GEMMs find it in the split cache by this key
                        dZ = DevMat(zhi, zhi.data_ptr(), rows, Nout, dY.ld)
                        self._split_cache[(dZ.ptr, dZ.rows, dZ.cols, dZ.ld)] = (zhi, zlo, zex, zhi)
                    else:
                        dZ = dY if owned else self.new(rows, Nout)
                        amax = None
                        if self.gemm_mode == 4:
                            self._split_cache.pop((dZ.ptr, dZ.rows, dZ.cols, dZ.ld), None)     # dY's split (if any) is stale
                            t = self._zeros2()
                            self._set_amax(dZ, t)
                            amax = t.data_ptr()
                        _lib.call('ipavsr_dense_bwd_prep', dY.ptr, dY.ld, y.ptr, y.ld, dZ.ptr, dZ.ld,
                                  G((l, 'b')).ptr if l.b is not None else None, rows, Nout, ACT[l.nonlinearity.name], 0,
                                  amax, st)
                self._proj_bwd(run, l.input_layer, xin, dZ, ar.mat((l, 'W')), G((l, 'W')))
            elif isinstance(l, L.BatchNormLayer):
                dy = gsegs[0]
                x = run.vals[l.input_layer][0]
                F = x.cols
                save, m_total = run.saved[l]
                bst = torch.empty(2 * F, dtype=torch.float64, device=self.device)
                _lib.call('ipavsr_bn_bwd_stats', dy.ptr, dy.ld, x.ptr, x.ld, save.data_ptr(), save.data_ptr() + 4 * F,
                          bst.data_ptr(), x.rows, F, st)
                if self.world is not None:
                    torch.distributed.all_reduce(bst, group=self.world[2])
                dx = self.new(x.rows, F)
                _lib.call('ipavsr_bn_bwd', dy.ptr, dy.ld, x.ptr, x.ld, ar.mat((l, 'gamma')).ptr, save.data_ptr(),
                          save.data_ptr() + 4 * F, bst.data_ptr(), dx.ptr, dx.ld, G((l, 'beta')).ptr,
                          G((l, 'gamma')).ptr, x.rows, F, m_total, 0, st)
                if self.world is not None:      # beta/gamma grads are already global sums: undo the later all-reduce
                    for key in ((l, 'beta'), (l, 'gamma')):
                        G(key).torch_view().mul_(1.0 / self.world[1])
                self._pass_grad(run, l.input_layer, [dx])
            elif isinstance(l, L.DropoutLayer):
                if l in run.saved:
                    keeps, scale = run.saved[l]
                    outs = []
                    for g, k in zip(gsegs, keeps):
                        o = self.new(g.rows, g.cols)
                        _lib.call('ipavsr_dropout', g.ptr, g.ld, k.data_ptr(), o.ptr, o.ld, g.rows, g.cols, scale, st)
                        outs.append(o)
                    gsegs = outs
                self._pass_grad(run, l.input_layer, gsegs)
            elif isinstance(l, L.DeltaLayer):
                if self.requires_grad.get(l.input_layer, False):
                    g = gsegs[0]
                    F = g.cols // 3
                    tgt, acc = self._grad_target(run, l.input_layer, [DevMat(None, 0, g.rows, F, _ld8(F))])
                    _lib.call('ipavsr_delta_bwd', g.ptr, g.ld, tgt[0].ptr, tgt[0].ld, N, T, F, run.window, acc, st)
            elif isinstance(l, L.LSTMLayer):
                H = l.num_units
                mask, gates, cell, hprev = run.saved[l]
                if l in run.bwd_ready:
                    ev, dG = run.bwd_ready.pop(l)
                    torch.cuda.current_stream(self.device).wait_event(ev)
                else:
                    dG = self._lstm_bwd_launch(run, l, gsegs[0], st, None)
                if not l.learn_init:
                    G((l, 'cell_init')).torch_view().zero_()
                    G((l, 'hid_init')).torch_view().zero_()
                if l not in run.lstm_db_done:
                    _lib.call('ipavsr_colsum', dG.ptr, dG.ld, G((l, 'b')).ptr, N * T, 4 * H, 0, st)
                # dW_hid = hprev^T dG
                self.gemm(hprev, dG, G((l, 'W_hid')), H, 4 * H, N * T, transA=1)
                self._proj_bwd(run, l.input_layers[0], run.vals[l.input_layers[0]], dG, ar.mat((l, 'W_in')),
                               G((l, 'W_in')))
            elif isinstance(l, L.AdaptiveElemwiseSumLayer):
                g = gsegs[0]
                ins = [self._single(run.vals[i]) for i in l.input_layers]
                ptrs = (C.c_void_p * len(ins))(*[i.ptr for i in ins])
                lds = (C.c_int * len(ins))(*[i.ld for i in ins])
                _lib.call('ipavsr_adasum_bwd_coeff', g.ptr, g.ld, ptrs, lds, len(ins), G((l, 'coeffs')).ptr, g.rows,
                          g.cols, 0, st)
                coeffs = ar.mat((l, 'coeffs'))
                for k, i in enumerate(l.input_layers):
                    if not self.requires_grad.get(i, False):
                        continue
                    tgt, acc = self._grad_target(run, i, [g])
                    _lib.call('ipavsr_copy2d', g.ptr, g.ld, tgt[0].ptr, tgt[0].ld, g.rows, g.cols,
                              coeffs.ptr + 4 * k, acc, st)
            elif isinstance(l, L.ElemwiseSumLayer):
                for i in l.input_layers:
                    self._pass_grad(run, i, gsegs)
            elif isinstance(l, L.ConcatLayer):
                if l in run.cat:
                    g = gsegs[0]            # one buffer over the whole width: every LSTM reads its column slice
                    for i, off in zip(l.input_layers, self.cat_plan[l][0]):
                        self._pass_grad(run, i, [DevMat(g.t, g.ptr + 4 * off, g.rows, i.num_units, g.ld)])
                else:
                    k = 0
                    for i in l.input_layers:
                        n = len(run.vals[i])
                        self._pass_grad(run, i, gsegs[k:k + n])
                        k += n
            elif isinstance(l, L.SliceLayer):
                g = gsegs[0]
                x = run.vals[l.input_layer][0]
                if l.input_layer in run.grads:
                    tgt, acc = self._grad_target(run, l.input_layer, [x])
                else:
                    full = self.new(x.rows, x.cols, zero=True)
                    run.grads[l.input_layer] = ([full], True)
                    tgt, acc = [full], 1
                _lib.call('ipavsr_slice_last', g.ptr, g.ld, tgt[0].ptr, tgt[0].ld, N, T, x.cols, 1, acc, st)
            else:
                raise TypeError('unsupported layer type %s' % type(l).__name__)
            if run.use_branches:
                run.keep.append(run.grads.pop(l))      # read on another stream than it was allocated on: freed with the run
            else:
                del run.grads[l]
            self._release(run, in_layers, remaining)
            if self._ar_hi is not None and l in self._flush_lo:
                self._ar_flush(self._flush_lo[l])
        if run.use_branches:
            self._branch_leave(run)
            for b in sorted(run.branches_used):
                self._branch_join(run, b)
        self._join(run)
        if self._ar_hi is not None:
            self._ar_flush(0, final=True)
        self._cur_run = None

    def _ar_flush(self, lo, final=False):
        """All-reduce the finalised tail [lo, hi) of the gradient arena once it is a bucket's worth (or the walk is over)."""
        if lo >= self._ar_hi or (not final and self._ar_hi - lo < self.ar_bucket_floats):
            return
        if final:
            # the side-stream recurrences were joined: everything is ordered before this call on the main stream
            lo = 0
        run = self._cur_run
        if run is not None and run.use_branches and not final:
            # the range was finalised by kernels on several streams (branches, trunk); the collective is ordered behind
            # the CURRENT stream only, so that stream first waits for the others
            cur = torch.cuda.current_stream(self.device)
            for s in [run.main_stream] + [self._branch_streams[b] for b in sorted(run.branches_used)]:
                if s != cur:
                    ev = torch.cuda.Event()
                    ev.record(s)
                    cur.wait_event(ev)
        first = self._ar_hi == self.arena.flat.numel()       # the bucket that carries the arena's tail: the loss sums
        self._ar_works.append(torch.distributed.all_reduce(self.arena.grad[lo:self._ar_hi], group=self.world[2],
                                                           async_op=True))
        if first and not final:
            self._loss_early_copy(self._ar_works[-1])
        self._ar_hi = lo

    def _release(self, run, ins, remaining):
        for i in ins:
            remaining[i] -= 1
            if (remaining[i] == 0 and self.concurrent_lstm and isinstance(i, L.LSTMLayer) and i in run.grads
                    and i not in run.bwd_ready):
                side, sh = self._fork()
                dG = self._lstm_bwd_launch(run, i, run.grads[i][0][0], sh, side)
                ev = torch.cuda.Event()
                ev.record(side)
                run.bwd_ready[i] = (ev, dG)

    def _lstm_bwd_launch(self, run, l, dout, stream_handle, side):
        ar, lib = self.arena, self.lib
        N, T, H = run.N, run.T, l.num_units
        G = lambda key: ar.mat(key, 'grad')
        mask, gates, cell, hprev = run.saved[l]
        dG = self.new(N * T, 4 * H)
        peep = ar.mat((l, 'peep')).ptr if l.peepholes else None
        dpeep = G((l, 'peep')).ptr if l.peepholes else None
        nbytes = lib.ipavsr_lstm_workspace_bytes(N, T, H)
        if side is not None:
            ws = torch.empty((int(nbytes) + 3) // 4, dtype=torch.float32, device=self.device)
            run.keep.append(ws)
            run.keep.append(dout.t)
        else:
            ws = self._workspace(nbytes)
        clip = l.grad_clipping if l.grad_clipping else 0.0
        whid = ar.mat((l, 'W_hid'))
        steps = (self.gemm_mode == 4 and self.lstm_impl == 0 and H > 256 and 0.0 < clip < 16384.0 and
                 not lib.ipavsr_lstm_bwd_f16_supported(N, T, H, whid.ld, float(clip)) and
                 lib.ipavsr_lstm_steps_supported(N, T, H, whid.ld) and dG.ld == 4 * H)
        if steps:
            wh, wl, we = self._split16(whid)
            n = max(dG.rows * dG.ld, 8)
            ghi = torch.empty(n, dtype=torch.float16, device=self.device)
            glo = torch.empty(n, dtype=torch.float16, device=self.device)
            gex = self._zeros2()
            self._split_cache[(dG.ptr, dG.rows, dG.cols, dG.ld)] = (ghi, glo, gex, dG.t)
            self._split_by_storage.setdefault(id(dG.t), []).append((dG.ptr, dG.rows, dG.cols, dG.ld))
            sbytes = int(lib.ipavsr_lstm_steps_workspace_bytes(N, T, H))
            sws = torch.empty((sbytes + 3) // 4, dtype=torch.float32, device=self.device)
            run.keep.append(sws)
            act = run.plan.active_rows.ctypes.data_as(C.c_void_p) if run.plan is not None else None
            _lib.call('ipavsr_lstm_bwd_f16_steps', dout.ptr, whid.ptr, wh, wl, we, whid.ld, peep,
                      ar.mat((l, 'cell_init')).ptr, mask.data_ptr(), gates.ptr, cell.ptr, dG.ptr, dpeep,
                      G((l, 'cell_init')).ptr, G((l, 'hid_init')).ptr, N, T, H, dout.ld, 1 if l.backwards else 0, clip, 0,
                      ghi.data_ptr(), glo.data_ptr(), gex.data_ptr() + 4, act, sws.data_ptr(), sbytes, stream_handle)
        elif (self.gemm_mode == 4 and self.lstm_impl == 0 and
                lib.ipavsr_lstm_bwd_f16_supported(N, T, H, whid.ld, float(clip))):
            wh, wl, we = self._split16(whid)
            # by-products of the kernel: the bias gradient and the fp16 split of dgates (dense 4H-wide rows)
            ghi = glo = gex = None
            if dG.ld == 4 * H:
                n = max(dG.rows * dG.ld, 8)
                ghi = torch.empty(n, dtype=torch.float16, device=self.device)
                glo = torch.empty(n, dtype=torch.float16, device=self.device)
                gex = self._zeros2()
                self._split_cache[(dG.ptr, dG.rows, dG.cols, dG.ld)] = (ghi, glo, gex, dG.t)
                self._split_by_storage.setdefault(id(dG.t), []).append((dG.ptr, dG.rows, dG.cols, dG.ld))
            _lib.call('ipavsr_lstm_bwd_f16', dout.ptr, whid.ptr, wh, wl, we, whid.ld, peep, ar.mat((l, 'cell_init')).ptr,
                      mask.data_ptr(), gates.ptr, cell.ptr, dG.ptr, dpeep, G((l, 'cell_init')).ptr,
                      G((l, 'hid_init')).ptr, N, T, H, dout.ld, 1 if l.backwards else 0, clip, 0,
                      G((l, 'b')).ptr, ghi.data_ptr() if ghi is not None else None,
                      glo.data_ptr() if glo is not None else None, gex.data_ptr() + 4 if gex is not None else None,
                      ws.data_ptr(), int(nbytes), stream_handle)
            run.lstm_db_done.add(l)
        else:
            _lib.call('ipavsr_lstm_bwd', dout.ptr, whid.ptr, peep, ar.mat((l, 'cell_init')).ptr,
                      mask.data_ptr(), gates.ptr, cell.ptr, dG.ptr, dpeep, G((l, 'cell_init')).ptr,
                      G((l, 'hid_init')).ptr, N, T, H, dout.ld, 1 if l.backwards else 0, clip, 0,
                      self.lstm_impl, ws.data_ptr(), int(nbytes), stream_handle)
        return dG

    def _proj_bwd(self, run, in_layer, xsegs, dZ, W, dW):
        """dW[rows of seg] = seg^T dZ ;  d(seg) (+)= dZ W[rows of seg]^T."""
        need_dx = self.requires_grad.get(in_layer, False)
        tgt = acc = None
        if need_dx:
            tgt, acc = self._grad_target(run, in_layer, xsegs)
        k0 = 0
        for i, a in enumerate(xsegs):
            if a.chunks and self.gemm_mode == 4 and len(xsegs) == 1:
                # x was split chunk by chunk (one scale each): the weight gradient accumulates over the row chunks
                self._split16(dZ)
                for ci, (r0, n, _) in enumerate(a.chunks):
                    self.gemm(a.row_slice(r0, n), dZ.row_slice(r0, n), dW, a.cols, dZ.cols, n, transA=1,
                              accumulate=1 if ci > 0 else 0)
                k0 += a.cols
                continue
            self.gemm(a, dZ, dW.row_slice(k0, a.cols), a.cols, dZ.cols, a.rows, transA=1)
            if need_dx:
                # the epilogue leaves |dX|max behind: the bound the next layer's fused backward prep scales its split with
                self.gemm(dZ, W.row_slice(k0, a.cols), tgt[i], a.rows, a.cols, dZ.cols, transB=1, accumulate=acc,
                          emit_split=(self.gemm_mode == 4))
            k0 += a.cols

    # ------------------------------------------------------------------------------------------------
    # loss + step
    # ------------------------------------------------------------------------------------------------
    def loss_and_backward(self, run, probs, loss, y, mask, count=None):
        """Runs the loss kernel (writes loss sum into the gradient arena tail) and the full backward."""
        self._pin_stream(True)
        try:
            return self._loss_and_backward(run, probs, loss, y, mask, count)
        finally:
            if run.use_branches and run.main_stream is not None:
                torch.cuda.set_stream(run.main_stream)      # also when the walk was left by an exception
            self._cur_run = None
            self._pin_stream(False)

    def _loss_and_backward(self, run, probs, loss, y, mask, count=None):
        st, ar = self.stream, self.arena
        plan = run.plan
        self._loss_ready = None
        # a packed / length-sorted run computes the loss in its own utterance order: targets follow, the mask is the plan's
        p = run.out_sorted[0] if plan is not None else probs[0]
        tail = ar.grad.data_ptr() + 4 * ar.tail
        ar.grad[ar.tail: ar.tail + 4].zero_()
        if self.world is not None:
            ar.grad[ar.tail + 2: ar.tail + 3].fill_(1.0)     # sums to world_size in the gradient all-reduce
        dlogits = self.new(p.rows, p.cols)
        if loss == 'squared_error':
            yd = None
        else:
            if run.targets_dev is None:
                self._stage_targets(run, y)
            yd = run.targets_dev
        count_dev = None
        if loss == 'temporal_softmax':
            if plan is not None:
                if count is None and not (isinstance(mask, torch.Tensor) and mask.is_cuda):
                    count = float(np.asarray(mask).sum())
                if count is None:
                    count = float(plan.M)
                md = plan.mask
            else:
                md = mask if isinstance(mask, torch.Tensor) and mask.is_cuda else self._upload(mask, 'mask')
            if count is None:
                count = float(np.asarray(mask).sum()) if not isinstance(mask, torch.Tensor) else None
            if count is None or self.world is not None:
                cnt = md.sum(dtype=torch.float32).reshape(1)
                if self.world is not None:
                    torch.distributed.all_reduce(cnt, group=self.world[2])
                ar.grad[ar.tail + 1: ar.tail + 2].copy_(cnt)
                count_dev, inv = tail + 4, 1.0
            else:
                ar.grad[ar.tail + 1: ar.tail + 2].fill_(count)
                inv = 1.0 / count
            _lib.call('ipavsr_temporal_softmax_loss', p.ptr, p.ld, yd.data_ptr(), md.data_ptr(), tail, dlogits.ptr,
                      dlogits.ld, p.rows, p.cols, inv, count_dev, st)
        elif loss == 'categorical_crossentropy':
            n_glob = p.rows * (self.world[1] if self.world is not None else 1)
            ar.grad[ar.tail + 1: ar.tail + 2].fill_(float(n_glob))
            _lib.call('ipavsr_categorical_crossentropy', p.ptr, p.ld, yd.data_ptr(), tail, dlogits.ptr, dlogits.ld,
                      p.rows, p.cols, 1.0 / n_glob, None, st)
        elif loss == 'squared_error':
            # T.mean(squared_error(pred, target)): mean over every element of the GLOBAL batch; dlogits is d/d(output)
            td = self._target_matrix(y, p)
            if plan is not None:
                td = self._gather(td.ptr, td.ld, td.rows, td.cols, plan.perm if td.rows == plan.N * plan.T else plan.order)
            n_glob = p.rows * p.cols * (self.world[1] if self.world is not None else 1)
            ar.grad[ar.tail + 1: ar.tail + 2].fill_(float(n_glob))
            _lib.call('ipavsr_squared_error', p.ptr, p.ld, td.ptr, td.ld, tail, dlogits.ptr, dlogits.ld, p.rows, p.cols,
                      1.0 / n_glob, st)
            if self.world is None:
                self._loss_early_copy()
            self.backward(run, dlogits, softmax_head=False)
            return
        else:
            raise ValueError('unknown loss %r' % (loss,))
        if self.world is None:
            self._loss_early_copy()
        self.backward(run, dlogits)

    def _target_matrix(self, y, p):
        """Regression targets (rows, F) or (N, T, F), host or device -> device matrix with the rows of the prediction."""
        if isinstance(y, DevMat):
            return y
        t = y if isinstance(y, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(y).astype(np.float32, copy=False)))
        if t.shape[-1] != p.cols or t.numel() != p.rows * p.cols:
            raise ValueError('targets of shape %r do not match predictions (%d, %d)' % (tuple(t.shape), p.rows, p.cols))
        return self._upload(t.reshape(p.rows, 1, p.cols), 'float')

    def _l2_tables(self, coef):
        """Per-tensor penalty coefficients over the arena: `coef` for tensors whose parameters are all regularizable."""
        ar = self.arena
        if getattr(self, '_l2_cache', None) is not None and self._l2_cache[0] == coef:
            return self._l2_cache[1], self._l2_cache[2]
        reg = {}
        for prm in L.get_all_params(self.out):
            if prm not in ar.bind or ar.bind[prm][1] == 'aux':
                continue
            k = ar.tensor_of(prm)
            reg[k] = reg.get(k, True) and ('regularizable' in prm.tags)
        cs = np.zeros(len(ar.order) + 1, dtype=np.float32)
        for i, k in enumerate(ar.order):
            cs[i] = coef if reg.get(k, False) else 0.0
        ids = np.zeros(ar.n // SEG + 1, dtype=np.int32)
        ids[:len(ar.seg_host)] = ar.seg_host
        ids[len(ar.seg_host):] = len(ar.order)            # the loss tail: coefficient 0
        seg_c = torch.from_numpy(cs).to(self.device)
        seg_i = torch.from_numpy(ids).to(self.device)
        self._l2_cache = (coef, seg_c, seg_i)
        return seg_c, seg_i

    def l2_penalty(self, coef, grads=True, loss_buf=None, count=None):
        """`+ coef * regularize_network_params(net, l2)`: adds 2 coef W to the gradients of the regularizable tensors and
        coef * sum W^2 to the loss (scaled by the loss normaliser, which read_loss divides by).  Data-parallel: rank 0 alone
        adds it, the gradient all-reduce (SUM) distributes it."""
        if self.world is not None and self.world[0] != 0:
            return
        ar = self.arena
        seg_c, seg_i = self._l2_tables(float(coef))
        if count is None:
            count = float(ar.grad[ar.tail + 1].item()) if loss_buf is None else 1.0
        tail = ar.grad.data_ptr() + 4 * ar.tail if loss_buf is None else loss_buf
        _lib.call('ipavsr_l2_penalty', ar.flat.data_ptr(), ar.grad.data_ptr() if grads else None, ar.n, seg_c.data_ptr(),
                  seg_i.data_ptr(), tail, float(count), self.stream)

    def loss_only(self, probs, loss, y, mask, l2=0.0):
        st = self.stream
        p = probs[0]
        buf = torch.zeros(4, dtype=torch.float32, device=self.device)
        if loss == 'squared_error':
            td = self._target_matrix(y, p)
            _lib.call('ipavsr_squared_error', p.ptr, p.ld, td.ptr, td.ld, buf.data_ptr(), None, 0, p.rows, p.cols, 1.0, st)
            both = torch.stack([buf[0], torch.tensor(float(p.rows * p.cols), dtype=torch.float32, device=self.device)])
            if self.world is not None:
                torch.distributed.all_reduce(both, group=self.world[2])
            h = both.cpu().numpy()
            val = np.float32(h[0] / h[1])
            if l2:
                pen = torch.zeros(1, dtype=torch.float32, device=self.device)
                seg_c, seg_i = self._l2_tables(float(l2))
                _lib.call('ipavsr_l2_penalty', self.arena.flat.data_ptr(), None, self.arena.n, seg_c.data_ptr(),
                          seg_i.data_ptr(), pen.data_ptr(), 1.0, st)
                val = np.float32(val + np.float32(pen.item()))
            return val
        yd = self._upload(y, 'int')
        if loss == 'temporal_softmax':
            md = self._upload(mask, 'mask')
            cnt = md.sum(dtype=torch.float32).reshape(1)
            _lib.call('ipavsr_temporal_softmax_loss', p.ptr, p.ld, yd.data_ptr(), md.data_ptr(), buf.data_ptr(), None,
                      0, p.rows, p.cols, 1.0, None, st)
            both = torch.stack([buf[0], cnt[0]])
        else:
            _lib.call('ipavsr_categorical_crossentropy', p.ptr, p.ld, yd.data_ptr(), buf.data_ptr(), None, 0, p.rows,
                      p.cols, 1.0, None, st)
            both = torch.stack([buf[0], torch.tensor(float(p.rows), dtype=torch.float32, device=self.device)])
        if self.world is not None:
            torch.distributed.all_reduce(both, group=self.world[2])
        h = both.cpu().numpy()
        return np.float32(h[0] / h[1])

    # ------------------------------------------------------------------------------------------------
    # small batches: the whole forward + loss + backward as ONE CUDA graph launch
    # ------------------------------------------------------------------------------------------------
    def graph_eligible(self, inputs, y, deterministic, l2):
        """The reference's own batches (26 utterances, `avletters/trimodal.py:356-359`; 10, `oulu/bimodal.py:360`) are
        launch-bound: ~140 launches whose enqueue takes the host longer than the device needs to run them.  Such a step is
        captured once per input shape (static input buffers, everything else allocated inside the capture) and replayed;
        the optimiser update stays outside the graph (its bias-correction scalar changes every step).  Not for: batches large
        enough to pack (the plan changes the shapes per batch), random dropout (the counter would be baked in), data
        parallelism, L2 penalties, derived or host-chunked inputs."""
        if self.graph_mode == 'off' or self._graph_failed or self.world is not None or l2 or len(self.mask_layers) > 1:
            return False
        if not deterministic and any(isinstance(l, L.DropoutLayer) for l in self.layers):
            return False
        rows = None
        for l in self.input_layers:
            a = inputs[l]
            if isinstance(a, Derived) or not hasattr(a, 'shape') or len(a.shape) not in (2, 3):
                return False
            rows = int(a.shape[0]) * int(a.shape[1])
        if rows is None or rows >= 2048 or self.packed_mode == 'force':
            return False
        return y is not None and hasattr(y, 'shape')

    def _to_static(self, buf, a):
        if isinstance(a, torch.Tensor):
            buf.copy_(a.to(buf.dtype) if a.dtype != buf.dtype else a, non_blocking=True)
        else:
            buf.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(a).astype(
                {torch.float32: np.float32, torch.uint8: np.uint8, torch.int32: np.int32}[buf.dtype], copy=False))),
                non_blocking=True)

    def graph_step(self, inputs, window, y, mask_layer, loss, deterministic):
        """forward + loss + backward of one small batch by replaying its CUDA graph (captured on first use per shape).
        Returns False when the capture failed (the caller then runs the eager path)."""
        key = (tuple((id(l), tuple(int(d) for d in inputs[l].shape)) for l in self.input_layers), int(window or 0), loss,
               tuple(int(d) for d in y.shape), bool(deterministic))
        ent = self._graphs.get(key)
        if ent is None:
            bufs = {}
            for l in self.input_layers:
                dt = torch.uint8 if l in self.mask_layers else torch.float32
                bufs[l] = torch.empty(tuple(int(d) for d in inputs[l].shape), dtype=dt, device=self.device)
            ybuf = torch.empty(tuple(int(d) for d in y.shape), dtype=torch.int32, device=self.device)
            for l in self.input_layers:
                self._to_static(bufs[l], inputs[l])
            self._to_static(ybuf, y)
            mbuf = bufs[mask_layer] if mask_layer is not None else None

            def body():
                run, out = self.forward(bufs, window, deterministic, train=True)
                self.loss_and_backward(run, out, loss, ybuf, mbuf, count=None)

            import gc
            err = None
            for attempt in range(2):
                try:
                    # one eager pass first: lazily created state (fp16 arenas, workspaces, constants, function attributes)
                    # must exist before the capture and live outside the graph's memory pool
                    body()
                    self.arena.split_dirty = True    # the captured step refreshes the operand split of the parameters
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    n0 = int(self.lib.ipavsr_launch_count())
                    # no garbage collection inside the capture: the finaliser of an old step's buffers or events running in
                    # the middle of it is the one thing here that is not under this function's control (a capture of the
                    # multi-stream step was seen invalidated once in ~5 full test runs; the second attempt is for that)
                    gc.collect()
                    gc_on = gc.isenabled()
                    gc.disable()
                    try:
                        with torch.cuda.graph(g):
                            body()
                    finally:
                        if gc_on:
                            gc.enable()
                    ent = (g, bufs, ybuf, int(self.lib.ipavsr_launch_count()) - n0)
                    self._graphs[key] = ent
                    if len(self._graphs) > 8:
                        self._graphs.pop(next(iter(self._graphs)))
                    err = None
                    break
                except Exception as e:
                    err = e
                    self._st_pin = None
                    try:
                        torch.cuda.synchronize(self.device)
                    except Exception:
                        pass
            if err is not None:                      # capture not possible here: stay on the eager path for good
                self._graph_failed = True
                import warnings
                warnings.warn('ipavsr_b200: CUDA-graph capture of the small-batch step failed (%s); running eagerly' % (err,))
                return False
            # the eager pass and the capture left valid gradients of THIS batch in the arena only via the eager pass;
            # replay once so that the state is exactly what a replayed step leaves
        g, bufs, ybuf, nlaunch = ent
        for l in self.input_layers:
            self._to_static(bufs[l], inputs[l])
        self._to_static(ybuf, y)
        g.replay()
        self.lib.ipavsr_launch_count_add(nlaunch)      # the kernels of this repository inside the replayed graph
        return True

    def allreduce_grads(self):
        if self.world is None:
            return
        if self._ar_works or self._ar_hi == 0:
            for w in self._ar_works:          # issued bucket by bucket during the backward walk
                w.wait()
            self._ar_works, self._ar_hi = [], None
        else:
            torch.distributed.all_reduce(self.arena.grad, group=self.world[2])

    def _loss_early_copy(self, work=None):
        """[loss sum, normaliser, rank count] -> pinned host memory, ordered behind the current stream (and behind the NCCL
        work that summed them over the ranks); read_loss() then waits for this copy only."""
        if not (self.early_loss and self._early_loss_ok) or torch.cuda.is_current_stream_capturing():
            return
        ar = self.arena
        if self._loss_stream is None:
            self._loss_stream = torch.cuda.Stream(device=self.device, priority=-1)
            self._loss_host = torch.empty(4, dtype=torch.float32).pin_memory()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        ls = self._loss_stream
        ls.wait_event(ev)
        with torch.cuda.stream(ls):
            if work is not None:
                work.wait()
            if os.environ.get('IPAVSR_EARLY_LOSS_DMA', '0') == '1':
                self._loss_host.copy_(ar.grad[ar.tail: ar.tail + 4], non_blocking=True)
            else:
                # written by a kernel straight into the (device-mapped) pinned buffer: a D2H copy would queue on a copy
                # engine behind the bulk upload of the next batch
                _lib.call('ipavsr_copy2d', ar.grad.data_ptr() + 4 * ar.tail, 4, self._loss_host.data_ptr(), 4, 1, 4, None, 0,
                          C.c_void_p(ls.cuda_stream))
            done = torch.cuda.Event()
            done.record(ls)
        self._loss_ready = done

    def read_loss(self):
        ar = self.arena
        ev, self._loss_ready = self._loss_ready, None
        if ev is not None:
            ev.synchronize()
            h = self._loss_host.numpy()[:3].copy()
        else:
            h = ar.grad[ar.tail: ar.tail + 3].cpu().numpy()
        w = h[2] if h[2] > 0 else 1.0      # the (already global) count was summed world_size times
        return np.float32(h[0] / (h[1] / w))

    def optim_step(self, kind, lr, params=None, lr_map=None, **hp):
        ar, st = self.arena, self.stream
        ar.split_dirty = True
        ar.touch_all()
        self._split_cache = {}
        self._split_by_storage = {}
        self._amax = {}
        n = ar.n
        seg_lr = seg_id = None
        if lr_map is not None:
            seg_lr, seg_id = self._lr_tables(lr_map, params)
        if kind == 'adam':
            b1, b2, eps = np.float32(hp.get('beta1', 0.9)), np.float32(hp.get('beta2', 0.999)), hp.get('epsilon', 1e-8)
            one = np.float32(1)
            self.step_t = np.float32(self.step_t + one)
            scalar = float(np.sqrt(one - b2 ** self.step_t) / (one - b1 ** self.step_t))
            _lib.call('ipavsr_optim_step', OPT['adam'], ar.flat.data_ptr(), ar.grad.data_ptr(),
                      ar.opt_state('m').data_ptr(), ar.opt_state('v').data_ptr(), n, float(lr),
                      seg_lr, seg_id, scalar, float(b1), float(b2), float(eps), 1.0, st)
        elif kind == 'adadelta':
            _lib.call('ipavsr_optim_step', OPT['adadelta'], ar.flat.data_ptr(), ar.grad.data_ptr(),
                      ar.opt_state('acc').data_ptr(), ar.opt_state('dacc').data_ptr(), n, float(lr), seg_lr, seg_id,
                      1.0, float(hp.get('rho', 0.95)), 0.0, float(hp.get('epsilon', 1e-6)), 1.0, st)
        elif kind == 'sgd':
            _lib.call('ipavsr_optim_step', OPT['sgd'], ar.flat.data_ptr(), ar.grad.data_ptr(), None, None, n,
                      float(lr), seg_lr, seg_id, 1.0, 0.0, 0.0, 0.0, 1.0, st)
        elif kind in ('momentum', 'nesterov'):
            _lib.call('ipavsr_optim_step', OPT[kind], ar.flat.data_ptr(), ar.grad.data_ptr(),
                      ar.opt_state('vel').data_ptr(), None, n, float(lr), seg_lr, seg_id, 1.0,
                      float(hp.get('momentum', 0.9)), 0.0, 0.0, 1.0, st)
        else:
            raise ValueError('unknown update rule %r' % (kind,))

    def _lr_tables(self, lr_map, params):
        """Per-tensor learning rates (adam_vlr, custom/updates.py:83): a device table indexed by 256-float block."""
        ar = self.arena
        key = tuple(sorted((p.name, float(v.get_value() if hasattr(v, 'get_value') else v)) for p, v in lr_map.items()))
        if self._lr_cache is not None and self._lr_cache[0] == key:
            return self._lr_cache[1], self._lr_cache[2]
        per_tensor = {}
        for p, v in lr_map.items():
            v = v.get_value() if hasattr(v, 'get_value') else v
            k = ar.tensor_of(p)
            if k in per_tensor and per_tensor[k] != float(v):
                raise ValueError('parameters sharing the device tensor %r need one learning rate' % (k,))
            per_tensor[k] = float(v)
        # the four gate matrices of an LSTM (and the three peephole vectors) live in one device tensor: updating only
        # some of them is not expressible with a per-tensor rate
        for p, (k, idx) in ar.bind.items():
            if idx != 'aux' and k in per_tensor and p not in lr_map and 'trainable' in p.tags and per_tensor[k] != 0.0:
                raise ValueError('parameter %s shares the device tensor of updated parameters but is not in the update list'
                                 % (p.name,))
        ids = np.zeros(ar.n // SEG + 1, dtype=np.int32)
        lrs = np.zeros(len(ar.order) + 1, dtype=np.float32)
        for i, k in enumerate(ar.order):
            off, rows, cols, ld = ar.tensors[k]
            nblk = (rows * ld + SEG - 1) // SEG
            ids[off // SEG: off // SEG + nblk] = i
            lrs[i] = per_tensor.get(k, 0.0)
        seg_id = torch.from_numpy(ids).to(self.device)
        seg_lr = torch.from_numpy(lrs).to(self.device)
        self._lr_cache = (key, seg_lr.data_ptr(), seg_id.data_ptr(), seg_lr, seg_id)
        return seg_lr.data_ptr(), seg_id.data_ptr()

    def param_grads(self, params=None):
        """Host copies of the gradients in Lasagne parameter order (tests)."""
        params = params if params is not None else L.get_all_params(self.out, trainable=True)
        return [self.arena.read(p, 'grad') for p in params]


def get_engine(output_layer, **kw):
    """One engine (one parameter arena) per output layer, shared by every compiled function."""
    eng = getattr(output_layer, '_ipavsr_engine', None)
    if eng is None:
        eng = Engine(output_layer, **kw)
        output_layer._ipavsr_engine = eng
    else:
        have = {'gemm_mode': [k for k, v in GEMM_MODES.items() if v == eng.gemm_mode][0], 'lstm_impl': eng.lstm_impl,
                'delta_exact': eng.delta_exact, 'packed': eng.packed_mode, 'device': str(eng.device)}
        for k, v in kw.items():
            if v is None:
                continue
            if k not in have:
                raise TypeError('unknown engine option %r' % (k,))
            ok = str(torch.device(v)) == have[k] if k == 'device' else (
                int(v) == have[k] if k in ('lstm_impl', 'delta_exact') else v == have[k])
            if not ok:
                raise ValueError('the engine of this network already exists with %s=%r; %r was asked for'
                                 % (k, have[k], v))
    return eng
