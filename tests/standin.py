"""Cheap deterministic stand-in for a Keras model (test infrastructure).

``predict`` maps float32 tiles [B,P,P,3] to a 2-channel 'softmax' [B,P,P,2] using only IEEE-exact elementwise
float32 operations (+ - * /), so results are bit-reproducible across numpy versions and machines, and it is
orientation-sensitive (position ramps), so any TTA / transpose mistake changes the output.
"""
import numpy as np


class StandInModel:
    def __init__(self, a=2.0, b=1.0, c=0.5):
        self.a, self.b, self.c = np.float32(a), np.float32(b), np.float32(c)

    def predict(self, x, batch_size=None, verbose=0, steps=None):
        x = np.asarray(x, dtype=np.float32)
        B, P = x.shape[0], x.shape[1]
        ri = (np.arange(P, dtype=np.float32) / np.float32(P))[None, :, None]
        rj = (np.arange(P, dtype=np.float32) / np.float32(P))[None, None, :]
        t = self.a * x[..., 0] - self.b * x[..., 1] * ri + self.c * x[..., 2] * rj + (ri - rj)
        p1 = np.float32(0.5) + np.float32(0.5) * t / (np.float32(1.0) + np.abs(t))
        return np.stack([np.float32(1.0) - p1, p1], axis=-1).astype(np.float32)
