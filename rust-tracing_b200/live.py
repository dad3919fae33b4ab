"""Progressive passes and checkpoint / resume.

The reference's live mode (renderer.rs:77-137) renders ONE sample of every pixel per UI tick and folds it into a
running mean `avg += (new - avg) / n` (renderer.rs:114). With SUM framebuffers and the sample-range render call that
is simply "accumulate sample index k, show sum / count": the mean after n passes is identical, and stopping /
resuming is saving the SUM buffer plus the next sample index (the reference cannot persist a render).
The window toolkit (fltk + pixels) is out of scope; `frames()` yields what the window would show.
"""
import numpy as np


class ProgressiveRender:
    def __init__(self, ctx, dscene, cam, seed=0, passes_per_tick=1):
        import torch
        self.ctx, self.dscene, self.cam, self.seed = ctx, dscene, cam, seed
        self.passes_per_tick = passes_per_tick
        h, w = cam.shape
        self.fb = torch.zeros((h, w, 4), dtype=torch.float32, device=f"cuda:{ctx.device_id}")
        self.next_sample = 0

    def tick(self):
        """One UI tick of renderer.rs:100-131: render the next pass(es) into the SUM buffer."""
        import torch
        stream = torch.cuda.current_stream(self.fb.device).cuda_stream
        self.ctx.render_accumulate(self.dscene, self.cam, self.next_sample, self.passes_per_tick, self.seed,
                                   self.fb.data_ptr(), stream)
        self.next_sample += self.passes_per_tick
        return self.next_sample

    def mean(self):
        """Running mean image (H, W, 3) float32 — what renderer.rs:114 keeps in raw_pixels."""
        f = self.fb.cpu().numpy()
        return f[..., :3] / np.maximum(f[..., 3:4], 1.0)

    def frame_rgb8(self):
        """The RGBA surface of renderer.rs:119-127 without the alpha channel (device-side color_to_rgb)."""
        h, w = self.cam.shape
        return self.ctx.finalize_rgb8(self.fb.data_ptr(), h * w, 0.0).reshape(h, w, 3)

    def frames(self, spp=None):
        """Like the window loop: renders while num_samples < spp. The reference starts num_samples at 1 and tests
        `<` (renderer.rs:98,104), so it shows spp-1 passes; same here."""
        spp = self.cam.samples_per_pixel if spp is None else spp
        num_samples = 1 + self.next_sample
        while num_samples < spp:
            self.tick()
            num_samples += self.passes_per_tick
            yield self.next_sample, self.frame_rgb8()

    def save(self, path):
        np.savez_compressed(path, sum_rgba=self.fb.cpu().numpy(), next_sample=self.next_sample, seed=self.seed,
                            shape=np.array(self.cam.shape))

    def load(self, path):
        import torch
        z = np.load(path)
        if tuple(z["shape"]) != self.cam.shape or int(z["seed"]) != self.seed:
            raise ValueError("checkpoint was made with another image size or seed")
        self.fb.copy_(torch.from_numpy(z["sum_rgba"]))
        self.next_sample = int(z["next_sample"])
        return self.next_sample
