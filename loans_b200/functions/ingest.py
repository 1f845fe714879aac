"""``FrameIngest`` / ``ingest_frames`` -- the loader's frame path on the device.

The reference's datasets hand the localizer ``resize_image(frame, image_size) / 255`` (reference
common/datasets/image_dataset.py:16-28, :98, :181): per sample, on the host, ``PIL.Image.resize(..., Image.LANCZOS)`` on the
uint8 HWC frame, then float32 CHW, then ``/ 255``.  Here a batch of decoded uint8 HWC frames already on the device becomes the
float32 NCHW batch in two integer kernels behind ``loans_stn_ingest_u8`` -- bit for bit Pillow's result (the coefficient
tables are computed on the host exactly as Pillow computes them).  No CPU fallback.
"""
import torch

from loans_b200 import _lib
from loans_b200.functions.spatial_transformer import InvalidType, _expect, _need_cuda, _on_device, _ptr, _stream


class FrameIngest(object):
    """Prepared ingest for frames of one size: ``FrameIngest(batch, (H, W), image_size)(frames_u8) -> float32 (B,3,oH,oW)``.

    The workspace (coefficient tables + the intermediate image of the two-pass resampling) is allocated and filled once;
    calls only launch kernels (and can be captured into a CUDA graph).  ``image_size=None``: no resampling, ``/ 255`` only.
    """

    def __init__(self, batch, frame_size, image_size=None, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("FrameIngest needs a CUDA device (loans_b200 has no CPU fallback)")
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.b, (self.h, self.w) = int(batch), (int(frame_size[0]), int(frame_size[1]))
        self.oh, self.ow = (self.h, self.w) if image_size is None else (int(image_size[0]), int(image_size[1]))
        L = _lib.lib()
        nbytes = L.loans_stn_ingest_workspace_bytes(self.b, self.h, self.w, self.oh, self.ow)
        if nbytes < 0:
            raise InvalidType("bad ingest dimensions %s -> %s" % ((self.h, self.w), (self.oh, self.ow)))
        self.workspace = torch.empty(int(nbytes), dtype=torch.uint8, device=self.dev)
        with _on_device(self.workspace):
            _lib.check(L.loans_stn_ingest_prepare(_ptr(self.workspace), self.h, self.w, self.oh, self.ow, _stream()),
                       "loans_stn_ingest_prepare")

    def __call__(self, frames, out=None):
        _need_cuda(frames)
        _expect(frames.dtype == torch.uint8, "frames.dtype == uint8 (got %s)" % frames.dtype)
        _expect(frames.dim() == 4 and frames.shape[3] == 3, "frames.shape == (B, H, W, 3) (got %s)" % (tuple(frames.shape),))
        b = frames.shape[0]
        _expect(b <= self.b and tuple(frames.shape[1:3]) == (self.h, self.w),
                "frames must be at most %d frames of %dx%d (got %s)" % (self.b, self.h, self.w, tuple(frames.shape)))
        frames = frames.contiguous()
        if out is None:
            out = torch.empty((b, 3, self.oh, self.ow), dtype=torch.float32, device=frames.device)
        with _on_device(frames):
            _lib.check(_lib.lib().loans_stn_ingest_u8(_ptr(frames), _ptr(out), _ptr(self.workspace), b, self.h, self.w, self.oh, self.ow,
                                                      _stream()), "loans_stn_ingest_u8")
        return out


def ingest_frames(frames, image_size=None):
    """uint8 (B,H,W,3) CUDA frames -> float32 (B,3,oH,oW) in [0,1]: ``resize_image(frame, image_size) / 255`` for a batch."""
    _need_cuda(frames)
    _expect(frames.dim() == 4, "frames.ndim == 4 (got %d)" % frames.dim())
    with _on_device(frames):
        return FrameIngest(frames.shape[0], frames.shape[1:3], image_size, device=frames.device)(frames)


__all__ = ["FrameIngest", "ingest_frames"]
