"""Synthetic "billiard ball" video: the workload of BASELINE.json's configs.  Discs of
distinct colours bounce elastically on a green table; ground-truth boxes (disc bbox +-2 px) stand in
for the YOLOv8 detector, which is out of scope and timed separately.  Deterministic in ``seed``.
"""
import numpy as np

_PALETTE = None


def _palette():
    global _PALETTE
    if _PALETTE is None:
        rng = np.random.default_rng(12345)
        cols = []
        for i in range(64):
            h = (i * 0.618033988749895) % 1.0
            s, v = 0.65 + 0.35 * ((i // 8) % 2), 0.95 - 0.25 * ((i // 16) % 2)
            k = int(h * 6)
            f = h * 6 - k
            p, q, t = v * (1 - s), v * (1 - f * s), v * (1 - (1 - f) * s)
            rgb = [(v, t, p), (q, v, p), (p, v, t), (p, q, v), (t, p, v), (v, p, q)][k % 6]
            cols.append([int(255 * c) for c in rgb])
        _PALETTE = np.array(cols, dtype=np.uint8)
        del rng
    return _PALETTE


class BilliardVideo:
    """frames(): iterator of uint8 RGB [H,W,3]; boxes(t): {obj_id: [x0,y0,x1,y1]} for frame t."""

    def __init__(self, num_objects=16, height=1024, width=1024, num_frames=300, seed=0, radius=None):
        self.B, self.H, self.W, self.T = num_objects, height, width, num_frames
        rng = np.random.default_rng(seed)
        self.radius = radius if radius is not None else max(4.0, 20.0 * height / 1024.0)
        g = int(np.ceil(np.sqrt(num_objects)))
        cx = (np.arange(num_objects) % g + 0.5) * (width / g)
        cy = (np.arange(num_objects) // g + 0.5) * (height / g)
        jit = min(width, height) / g * 0.2
        self.pos0 = np.stack([cx, cy], 1) + rng.uniform(-jit, jit, size=(num_objects, 2))
        self.vel = rng.uniform(-6, 6, size=(num_objects, 2)) * (height / 1024.0)
        self.noise_seed = int(rng.integers(0, 2 ** 31 - 1))
        self._traj = self._simulate()

    def _simulate(self):
        r = self.radius
        lo = np.array([r, r])
        hi = np.array([self.W - 1 - r, self.H - 1 - r])
        pos, vel = self.pos0.copy(), self.vel.copy()
        pos = np.clip(pos, lo, hi)
        traj = np.zeros((self.T, self.B, 2))
        for t in range(self.T):
            traj[t] = pos
            pos = pos + vel
            for d in range(2):
                under, over = pos[:, d] < lo[d], pos[:, d] > hi[d]
                pos[under, d] = 2 * lo[d] - pos[under, d]
                pos[over, d] = 2 * hi[d] - pos[over, d]
                vel[under | over, d] *= -1
        return traj

    def centers(self, t):
        return self._traj[t]

    def frame(self, t):
        rng = np.random.default_rng(self.noise_seed + t)
        img = np.empty((self.H, self.W, 3), dtype=np.int16)
        img[:] = np.array([30, 110, 60], dtype=np.int16)
        img += rng.integers(-4, 5, size=img.shape, dtype=np.int16)
        yy, xx = np.mgrid[0:self.H, 0:self.W]
        pal = _palette()
        for i, (cx, cy) in enumerate(self._traj[t]):
            x0, x1 = int(max(0, cx - self.radius - 1)), int(min(self.W, cx + self.radius + 2))
            y0, y1 = int(max(0, cy - self.radius - 1)), int(min(self.H, cy + self.radius + 2))
            m = (xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= self.radius ** 2
            img[y0:y1, x0:x1][m] = pal[i % 64].astype(np.int16)
        return np.clip(img, 0, 255).astype(np.uint8)

    def frames(self, start=0, stop=None):
        for t in range(start, self.T if stop is None else stop):
            yield self.frame(t)

    def boxes(self, t):
        r = self.radius + 2
        out = {}
        for i, (cx, cy) in enumerate(self._traj[t]):
            out[i] = [float(max(0, cx - r)), float(max(0, cy - r)), float(min(self.W - 1, cx + r)),
                      float(min(self.H - 1, cy + r))]
        return out


class GroundTruthDetector:
    """Stand-in for the YOLOv8 box-prompt detector on synthetic video ("ground-truth boxes =
    disc bbox +-2 px; detector time reported separately as n.a.").  Presents the injected-detector interface of
    ``VideoProcessor``: called with the frames that fall on the ``detect_interval`` grid, in stream order, and
    returns one detection list per frame (class = object id, as Det-SAM2 uses the class as obj_id)."""

    def __init__(self, video, detect_interval, first_frame=0, appear_at=None, repeat_class=None):
        self.video, self.interval, self.next_idx = video, detect_interval, first_frame
        self.appear_at = appear_at or {}  # obj_id -> first frame at which the detector reports it
        # obj_id -> (dx, dy): the detector reports a SECOND box of that class, shifted, after all first boxes (YOLO does
        # emit several instances of one class; Det-SAM2 then prompts the same obj_id twice on one frame)
        self.repeat_class = repeat_class or {}

    def __call__(self, frames_bgr):
        out = []
        for _ in frames_bgr:
            t = self.next_idx
            dets = [{"coordinates": np.asarray(b, np.float32), "class": np.asarray([float(oid)], np.float32),
                     "confidence": np.asarray([0.99], np.float32)}
                    for oid, b in self.video.boxes(t).items() if t >= self.appear_at.get(oid, 0)]
            for oid, (dx, dy) in self.repeat_class.items():
                if oid in self.video.boxes(t) and t >= self.appear_at.get(oid, 0):
                    b = np.asarray(self.video.boxes(t)[oid], np.float32) + np.asarray([dx, dy, dx, dy], np.float32)
                    dets.append({"coordinates": b, "class": np.asarray([float(oid)], np.float32),
                                 "confidence": np.asarray([0.95], np.float32)})
            out.append(dets)
            self.next_idx += self.interval
        return out
