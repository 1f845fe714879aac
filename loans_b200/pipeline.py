"""Host-buffer pipeline around the fused crop kernels: pinned host tensors in, pinned host tensors out.

A LoANs-style consumer that keeps its frames on the host (the reference's ``prepare_images`` round-trips the whole
batch through the host every step, sheep/sheep_localizer.py:72-82) pays PCIe for every step.  ``HostCropPipeline``
hides as much of that as the link allows: ``depth`` sets of device buffers, three CUDA streams (H2D, compute, D2H)
chained with events, so that the upload of step i+1 and the download of step i-1 run while step i computes --
PCIe is full duplex.  The compute stream calls the C ABI directly (``loans_stn_crop_fwd`` / ``loans_stn_crop_bwd``).

    pipe = HostCropPipeline(batch, channels, height, width, out_size, crops_per_frame=1, need_gx=True)
    for step in ...:
        pipe.submit(x_host, theta_host, gy_host, outputs)     # outputs: dict of pinned tensors y, grid, gtheta[, gx]
    pipe.drain()                                              # every submitted step's outputs are now on the host

All host tensors must be pinned (``tensor.pin_memory()``) for the copies to be asynchronous; they must stay alive and
unmodified until ``drain()`` (or until ``depth`` further submits have been made).
"""
import torch

from loans_b200 import _lib


class HostCropPipeline(object):
    def __init__(self, batch, channels, height, width, out_size, crops_per_frame=1, need_gx=True,
                 out_dtype=torch.float32, depth=2, device=None, uint8_frames=False):
        if not torch.cuda.is_available():
            raise RuntimeError("HostCropPipeline needs a CUDA device (loans_b200 has no CPU fallback)")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dims = (int(batch), int(crops_per_frame), int(channels), int(height), int(width), int(out_size[0]), int(out_size[1]))
        b, k, c, h, w, oh, ow = self.dims
        n = b * k
        self.need_gx = bool(need_gx)
        self.dt = out_dtype
        self.dt_code = _lib.BF16 if out_dtype == torch.bfloat16 else _lib.F32
        self.depth = int(depth)
        self.lib = _lib.lib()
        # uint8_frames: the host hands over DECODED frames, (B,H,W,3) uint8 as the loader reads them, and the `/ 255` float32
        # NCHW conversion (reference common/datasets/image_dataset.py:98) runs on the device (loans_stn_ingest_u8): a quarter
        # of the upload
        self.uint8_frames = bool(uint8_frames)
        if self.uint8_frames and int(channels) != 3:
            raise ValueError("uint8 frames are RGB (3 channels)")
        with torch.cuda.device(self.dev):
            self.s_in, self.s_run, self.s_out = (torch.cuda.Stream() for _ in range(3))
            self.slots = []
            if self.uint8_frames:
                from loans_b200.functions.ingest import FrameIngest
                self.ingest = FrameIngest(b, (h, w), None, device=self.dev)
            ydt = out_dtype
            self._in_fields = [("xu8", (b, h, w, 3), torch.uint8) if self.uint8_frames else ("x", (b, c, h, w), torch.float32),
                               ("theta", (n, 2, 3), torch.float32), ("gy", (n, c, oh, ow), ydt)]
            self._out_fields = [("y", (n, c, oh, ow), ydt), ("grid", (n, 2, oh, ow), torch.float32), ("gtheta", (n, 2, 3), torch.float32)]
            if need_gx:
                self._out_fields.append(("gx", (b, c, h, w), torch.float32))
            for _ in range(self.depth):
                # inputs and outputs of a slot live in ONE device allocation each, so that a step whose host tensors are packed the
                # same way (new_host_inputs / new_host_outputs) moves as one copy per direction instead of three and four
                s_in, in_views = self._packed(self._in_fields, self.dev, False)
                s_out, out_views = self._packed(self._out_fields, self.dev, False)
                slot = {"xu8": None, "x": None, "gx": None, "in_packed": s_in, "out_packed": s_out,
                        "ev_in": torch.cuda.Event(), "ev_run": torch.cuda.Event(), "ev_out": torch.cuda.Event(), "used": False}
                slot.update(in_views)
                slot.update(out_views)
                if self.uint8_frames:
                    slot["x"] = torch.empty((b, c, h, w), device=self.dev)
                self.slots.append(slot)
        self.step = 0
        self.h2d_bytes = (1 if self.uint8_frames else 4) * b * c * h * w + 4 * n * 6 + n * c * oh * ow * (2 if out_dtype == torch.bfloat16 else 4)
        self.d2h_bytes = n * c * oh * ow * (2 if out_dtype == torch.bfloat16 else 4) + 4 * (n * 2 * oh * ow + n * 6) \
            + (4 * b * c * h * w if need_gx else 0)

    @staticmethod
    def _packed(fields, device, pinned):
        """One uint8 allocation holding the fields back to back (256-byte aligned), and a dict of typed views into it."""
        offs, total = [], 0
        for _, shape, dt in fields:
            offs.append(total)
            nbytes = int(torch.Size(shape).numel()) * torch.empty((), dtype=dt).element_size()
            total += (nbytes + 255) // 256 * 256
        buf = torch.empty(total, dtype=torch.uint8, device=device)
        if pinned:
            buf = buf.pin_memory()
        views = {}
        for (name, shape, dt), off in zip(fields, offs):
            nbytes = int(torch.Size(shape).numel()) * torch.empty((), dtype=dt).element_size()
            views[name] = buf[off:off + nbytes].view(dt).view(shape)
        return buf, views

    def new_host_inputs(self):
        """Pinned host tensors for one step's inputs ('x', 'theta', 'gy'), views of ONE pinned buffer laid out like the device
        side: handed to submit() they upload as a single copy."""
        buf, views = self._packed(self._in_fields, "cpu", True)
        if self.uint8_frames:
            views["x"] = views.pop("xu8")
        views["_packed"] = buf
        return views

    def new_host_outputs(self):
        """Pinned host tensors for one step's results ('y', 'grid', 'gtheta'[, 'gx']), views of ONE pinned buffer: one download."""
        buf, views = self._packed(self._out_fields, "cpu", True)
        views["_packed"] = buf
        return views

    def submit(self, x_host, theta_host, gy_host, outputs, mask01=0.0, inputs=None):
        """Enqueue one fwd+bwd step.  ``outputs``: dict with pinned host tensors 'y', 'grid', 'gtheta' and, if need_gx, 'gx'.
        ``inputs`` / ``outputs`` made by new_host_inputs() / new_host_outputs() travel as ONE copy per direction (``inputs`` then
        replaces x_host / theta_host / gy_host, which may be None)."""
        b, k, c, h, w, oh, ow = self.dims
        n = b * k
        s = self.slots[self.step % self.depth]
        self.step += 1
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.s_in):
                if s["used"]:
                    self.s_in.wait_event(s["ev_out"])            # the slot's previous results have left the device
                if inputs is not None and inputs.get("_packed") is not None and inputs["_packed"].numel() == s["in_packed"].numel():
                    s["in_packed"].copy_(inputs["_packed"], non_blocking=True)        # x | theta | gy in one transfer
                else:
                    (s["xu8"] if self.uint8_frames else s["x"]).copy_(x_host, non_blocking=True)
                    s["theta"].copy_(theta_host, non_blocking=True)
                    s["gy"].copy_(gy_host, non_blocking=True)
                s["ev_in"].record(self.s_in)
            with torch.cuda.stream(self.s_run):
                self.s_run.wait_event(s["ev_in"])
                st = self.s_run.cuda_stream
                p = lambda t: None if t is None else t.data_ptr()          # noqa: E731
                if self.uint8_frames:
                    self.ingest(s["xu8"], out=s["x"])
                _lib.check(self.lib.loans_stn_crop_fwd(p(s["x"]), p(s["theta"]), float(mask01), p(s["y"]), p(s["grid"]),
                                                       n, k, c, h, w, oh, ow, self.dt_code, st), "loans_stn_crop_fwd")
                _lib.check(self.lib.loans_stn_crop_bwd(p(s["x"]), p(s["theta"]), float(mask01), p(s["gy"]), None, p(s["gtheta"]),
                                                       p(s["gx"]), None, n, k, c, h, w, oh, ow, self.dt_code, st),
                           "loans_stn_crop_bwd")
                s["ev_run"].record(self.s_run)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(s["ev_run"])
                if outputs.get("_packed") is not None and outputs["_packed"].numel() == s["out_packed"].numel():
                    outputs["_packed"].copy_(s["out_packed"], non_blocking=True)     # y | grid | gtheta [| gx] in one transfer
                else:
                    outputs["y"].copy_(s["y"], non_blocking=True)
                    outputs["grid"].copy_(s["grid"], non_blocking=True)
                    outputs["gtheta"].copy_(s["gtheta"], non_blocking=True)
                    if self.need_gx:
                        outputs["gx"].copy_(s["gx"], non_blocking=True)
                s["ev_out"].record(self.s_out)
            s["used"] = True

    def drain(self):
        self.s_out.synchronize()
        self.s_run.synchronize()
        self.s_in.synchronize()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.drain()

    def __del__(self):
        # the slot tensors are used on the side streams only: nothing may hand them back to the allocator while copies or
        # kernels are still in flight
        try:
            self.drain()
        except Exception:
            pass
