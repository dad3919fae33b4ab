"""Progressive passes and checkpoint / resume.

The reference's live mode (renderer.rs:77-137) renders ONE sample of every pixel per UI tick and folds it into a
running mean `avg += (new - avg) / n` (renderer.rs:114). With SUM framebuffers and the sample-range render call that
is simply "accumulate sample index k, show sum / count": the mean after n passes is identical, and stopping /
resuming is saving the SUM buffer plus the next sample index (the reference cannot persist a render).
The window toolkit (fltk + pixels) is out of scope; `frames()` yields what the window would show.
"""
import ctypes as C
import hashlib

import numpy as np


def render_identity(scene, cam, seed):
    """What a checkpoint belongs to: a digest of the flattened scene (every op word the device walks, materials and
    textures through the description arrays), the camera (rt_camera_desc, which carries max_depth) and the seed."""
    from . import api
    h = hashlib.sha256()
    if scene is not None and scene.desc is not None:
        h.update(np.ascontiguousarray(api.scene_ops(scene)["words"]).tobytes())
        d = scene.desc
        for ptr, n, size in ((d.textures, d.n_textures, None), (d.materials, d.n_materials, None)):
            if n > 0:
                h.update(C.string_at(ptr, n * C.sizeof(ptr._type_)))
    h.update(bytes(cam))
    h.update(int(seed).to_bytes(8, "little", signed=False))
    return h.hexdigest()


class ProgressiveRender:
    def __init__(self, ctx, dscene, cam, seed=0, passes_per_tick=1, scene=None):
        """`scene` (the host-side Scene the device scene was uploaded from) makes checkpoints carry the scene's
        identity; without it only camera, depth and seed are checked on resume."""
        import torch
        self.ctx, self.dscene, self.cam, self.seed = ctx, dscene, cam, seed
        self.passes_per_tick = passes_per_tick
        h, w = cam.shape
        self.fb = torch.zeros((h, w, 4), dtype=torch.float32, device=f"cuda:{ctx.device_id}")
        self.next_sample = 0
        self.identity = render_identity(scene, cam, seed)

    def _stream(self):
        import torch
        return torch.cuda.current_stream(self.fb.device).cuda_stream

    def tick(self, passes=None):
        """One UI tick of renderer.rs:100-131: render the next pass(es) into the SUM buffer."""
        passes = self.passes_per_tick if passes is None else passes
        self.ctx.render_accumulate(self.dscene, self.cam, self.next_sample, passes, self.seed, self.fb.data_ptr(), self._stream())
        self.next_sample += passes
        return self.next_sample

    def mean(self):
        """Running mean image (H, W, 3) float32 — what renderer.rs:114 keeps in raw_pixels."""
        f = self.fb.cpu().numpy()
        return f[..., :3] / np.maximum(f[..., 3:4], 1.0)

    def frame_rgb8(self):
        """The RGBA surface of renderer.rs:119-127 without the alpha channel (device-side color_to_rgb), on the stream
        the passes were rendered on."""
        h, w = self.cam.shape
        return self.ctx.finalize_rgb8(self.fb.data_ptr(), h * w, 0.0, self._stream()).reshape(h, w, 3)

    def frames(self, spp=None):
        """Like the window loop: renders while num_samples < spp. The reference starts num_samples at 1 and tests
        `<` (renderer.rs:98,104), so it shows spp-1 passes; same here - the last tick is shortened so that
        passes_per_tick > 1 never renders past it."""
        spp = self.cam.samples_per_pixel if spp is None else spp
        while 1 + self.next_sample < spp:
            self.tick(min(self.passes_per_tick, spp - 1 - self.next_sample))
            yield self.next_sample, self.frame_rgb8()

    def save(self, path):
        np.savez_compressed(path, sum_rgba=self.fb.cpu().numpy(), next_sample=self.next_sample, seed=self.seed,
                            shape=np.array(self.cam.shape), identity=self.identity)

    def load(self, path):
        import torch
        z = np.load(path)
        if tuple(z["shape"]) != self.cam.shape or int(z["seed"]) != self.seed:
            raise ValueError("checkpoint was made with another image size or seed")
        if "identity" not in z or str(z["identity"]) != self.identity:
            raise ValueError("checkpoint belongs to another scene, camera or max_depth; refusing to blend it into this render")
        self.fb.copy_(torch.from_numpy(z["sum_rgba"]))
        self.next_sample = int(z["next_sample"])
        return self.next_sample
