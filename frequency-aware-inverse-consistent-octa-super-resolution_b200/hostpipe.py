"""Host-buffer front end: run a device step over a batch that lives in pinned host memory.

The wavelet / SSIM path is batch-sharded -- every image plane is independent (pw/dwt/lowlevel.py:143 ``groups=C``) --
so a batch can be cut along dim 0 into chunks whose host->device copy, kernels and device->host copy overlap on
three CUDA streams (the two copy engines of a B200 move data in both directions at once).  With inputs and
results on the host, PCIe is the bound; overlapping the directions and hiding the kernels behind the copies is
what the end-to-end number of ``bench.py`` measures.

    pipe = HostPipeline(step, example_inputs=(x_host, g_host), chunks=4)
    rec_host, dx_host = pipe((x_host, g_host))

``step(*device_inputs) -> tuple of device tensors`` is called once per chunk (and captured into one CUDA graph per
chunk when ``graph=True``, so the per-chunk launch cost is a graph replay).  Outputs whose leading dimension is
the chunk's batch size are written to the matching rows of the pinned result; other outputs (e.g. a scalar
loss) come back stacked as ``(chunks, ...)`` and are the caller's to combine.
"""
import torch


def _chunk_bounds(n, chunks):
    base, rem = divmod(n, chunks)
    out, lo = [], 0
    for i in range(chunks):
        hi = lo + base + (1 if i < rem else 0)
        out.append((lo, hi))
        lo = hi
    return [b for b in out if b[1] > b[0]]


class HostPipeline(object):
    def __init__(self, step, example_inputs, chunks=4, graph=True, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("HostPipeline needs a CUDA device (there is no CPU fallback)")
        self.step = step
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        n = example_inputs[0].shape[0]
        self.bounds = _chunk_bounds(n, max(1, min(int(chunks), n)))
        self.n = n
        self.s_in, self.s_run, self.s_out = (torch.cuda.Stream(self.device) for _ in range(3))
        self.dev_in, self.dev_out, self.graphs = [], [], []
        self.ev_in = [torch.cuda.Event() for _ in self.bounds]
        self.ev_run = [torch.cuda.Event() for _ in self.bounds]
        self.ev_free = [torch.cuda.Event() for _ in self.bounds]   # chunk's device inputs may be overwritten
        self.ev_out = [torch.cuda.Event() for _ in self.bounds]    # chunk's device outputs have been copied out
        cur = torch.cuda.current_stream(self.device)
        self.s_run.wait_stream(cur)
        with torch.cuda.stream(self.s_run):
            for lo, hi in self.bounds:
                ins = tuple(torch.empty((hi - lo,) + tuple(t.shape[1:]), dtype=t.dtype, device=self.device)
                            .requires_grad_(bool(getattr(t, "requires_grad", False))) for t in example_inputs)
                with torch.no_grad():
                    for d, s in zip(ins, example_inputs):
                        d.copy_(s[lo:hi], non_blocking=True)
                for _ in range(2):   # warm-up: fills host-side caches (filter taps) before any capture
                    outs = self._call(ins)
                self.dev_in.append(ins)
                self.dev_out.append(outs)
            self.s_run.synchronize()
            if graph:
                for k, ins in enumerate(self.dev_in):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.s_run):
                        outs = self._call(ins)
                    self.graphs.append(g)
                    self.dev_out[k] = outs
        cur.wait_stream(self.s_run)
        self.host_out = None

    def _call(self, ins):
        for t in ins:
            if t.requires_grad:
                t.grad = None
        outs = self.step(*ins)
        if isinstance(outs, torch.Tensor):
            outs = (outs,)
        return tuple(o.detach() for o in outs if o is not None)

    def _alloc_host_out(self):
        outs = []
        for j, o in enumerate(self.dev_out[0]):
            per_sample = o.dim() > 0 and o.shape[0] == self.bounds[0][1] - self.bounds[0][0]
            shape = ((self.n,) + tuple(o.shape[1:])) if per_sample else ((len(self.bounds),) + tuple(o.shape))
            outs.append((torch.empty(shape, dtype=o.dtype).pin_memory(), per_sample))
        return outs

    def bytes_per_call(self, host_inputs):
        """(host->device, device->host) bytes one call moves."""
        if self.host_out is None:
            self.host_out = self._alloc_host_out()
        h2d = sum(t.numel() * t.element_size() for t in host_inputs)
        d2h = sum(t.numel() * t.element_size() for t, _ in self.host_out)
        return h2d, d2h

    def join(self):
        """Make the current stream wait for every result copy issued so far (device-side; does not block the host)."""
        torch.cuda.current_stream(self.device).wait_stream(self.s_out)

    def __call__(self, host_inputs, sync=True):
        """Copy in, run, copy out, chunk by chunk; returns the tuple of pinned result tensors.  They are valid after
        the call when ``sync`` is true.  With ``sync=False`` nothing waits: consecutive calls stream through the
        three queues back to back (call N+1's uploads overlap call N's downloads); use ``join()`` /
        ``s_out.synchronize()`` before reading the results, and note that the same pinned result buffers are reused by
        every call."""
        if self.host_out is None:
            self.host_out = self._alloc_host_out()
        cur = torch.cuda.current_stream(self.device)
        if sync:
            self.s_in.wait_stream(cur)   # inputs produced by work queued on the caller's stream
        for k, (lo, hi) in enumerate(self.bounds):
            with torch.cuda.stream(self.s_in):
                self.s_in.wait_event(self.ev_free[k])     # previous call's kernels are done with these buffers
                with torch.no_grad():
                    for d, s in zip(self.dev_in[k], host_inputs):
                        d.copy_(s[lo:hi], non_blocking=True)
                self.ev_in[k].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(self.ev_in[k])
                self.s_run.wait_event(self.ev_out[k])     # previous call's results have left the device
                if self.graphs:
                    self.graphs[k].replay()
                else:
                    self.dev_out[k] = self._call(self.dev_in[k])
                self.ev_run[k].record(self.s_run)
                self.ev_free[k].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_run[k])
                for (h, per_sample), o in zip(self.host_out, self.dev_out[k]):
                    (h[lo:hi] if per_sample else h[k]).copy_(o, non_blocking=True)
                self.ev_out[k].record(self.s_out)
        if sync:
            cur.wait_stream(self.s_out)
            self.s_out.synchronize()
        return tuple(h for h, _ in self.host_out)
