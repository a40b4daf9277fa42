"""Python restatement of the reference's addFrame() front-end (src/photobundle.cc:482-608:
visibility update by ZNCC, new-point creation at saliency local maxima, top-N selection,
descriptor extraction) used to check the C++ host class.  Test infrastructure only.
float32 arithmetic mirrors the C++ expression types (interp2 / ZnccPatch_ with T = float)."""
import numpy as np

f32 = np.float32


def interp2_u8(I, xf, yf, fill=f32(0.0)):
    rows, cols = I.shape
    max_cols, max_rows = cols - 1, rows - 1
    xi, yi = int(np.floor(xf)), int(np.floor(yf))
    xf = f32(xf - f32(xi)); yf = f32(yf - f32(yi))
    if 0 <= xi < max_cols and 0 <= yi < max_rows:
        wx = f32(1.0 - float(xf))
        a = f32(f32(f32(I[yi, xi]) * wx) + f32(f32(I[yi, xi + 1]) * xf))
        b = f32(f32(f32(I[yi + 1, xi]) * wx) + f32(f32(I[yi + 1, xi + 1]) * xf))
        return f32((1.0 - float(yf)) * float(a) + float(f32(yf * b)))
    if xi == max_cols and 0 <= yi < max_rows:
        return fill if xf > 0 else f32((1.0 - float(yf)) * float(I[yi, xi]) + float(yf) * float(I[yi + 1, xi]))
    if yi == max_rows and 0 <= xi < max_cols:
        return fill if yf > 0 else f32((1.0 - float(xf)) * float(I[yi, xi]) + float(xf) * float(I[yi, xi + 1]))
    if xi == max_cols and yi == max_rows:
        return fill if (xf > 0 or yf > 0) else f32(I[yi, xi])
    return fill


class Zncc:
    def __init__(self, I, px, py):
        x, y = f32(px), f32(py)
        d = np.array([interp2_u8(I, f32(f32(c) + x), f32(f32(r) + y)) for r in range(-2, 3) for c in range(-2, 3)], dtype=f32)
        s = f32(0)
        for v in d:
            s = f32(s + v)
        mean = f32(s / f32(25.0))
        d = (d - mean).astype(f32)
        ss = f32(0)
        for v in d:
            ss = f32(ss + f32(v * v))
        self.data, self.norm = d, f32(np.sqrt(ss))

    def score(self, o):
        d = f32(self.norm * o.norm)
        dot = f32(0)
        for a, b in zip(self.data, o.data):
            dot = f32(dot + f32(a * b))
        return f32(dot / d) if float(d) > 1e-6 else f32(-1.0)


def local_maxima(S, mask, n):
    """IsLocalMax_ (src/imgproc.h:175-212) for every pixel: unmasked, not below 0, and strictly above all neighbours of the
    (2n+1)^2 window (a tie loses); radius 0 accepts everything.  Pixels whose window leaves the image compare against +inf
    (the callers never ask for them)."""
    rows, cols = S.shape
    out = np.ones((rows, cols), dtype=bool)
    if n > 0:
        out &= mask & ~(S < 0)
        for r in range(-n, n + 1):
            for c in range(-n, n + 1):
                if r == 0 and c == 0:
                    continue
                sh = np.full((rows, cols), np.inf, dtype=f32)
                ys = slice(max(0, -r), rows - max(0, r)); yd = slice(max(0, r), rows - max(0, -r))
                xs = slice(max(0, -c), cols - max(0, c)); xd = slice(max(0, c), cols - max(0, -c))
                sh[ys, xs] = S[yd, xd]
                out &= ~(sh >= S)
    return out


def extract_patch(I, x, y, radius):
    rows, cols = I.shape
    mc, mr = cols - radius - 1, rows - radius - 1
    return np.array([float(I[max(radius, min(y + r, mr)), max(radius, min(x + c, mc))])
                     for r in range(-radius, radius + 1) for c in range(-radius, radius + 1)])


class RefFrontEnd:
    def __init__(self, rows, cols, K4, maxNumPoints=4096, slidingWindowSize=5, patchRadius=2, maskBlockRadius=1,
                 maxFrameDistance=1, minScore=0.75, minValidDepth=0.01, maxValidDepth=1000.0, nonMaxSuppRadius=1):
        self.rows, self.cols = rows, cols
        fx, fy, cx, cy = K4
        self.K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
        self.Kinv = np.linalg.inv(self.K)
        self.o = dict(maxNumPoints=maxNumPoints, patchRadius=patchRadius, maskBlockRadius=maskBlockRadius,
                      maxFrameDistance=maxFrameDistance, minScore=minScore, minValidDepth=minValidDepth,
                      maxValidDepth=maxValidDepth, nonMaxSuppRadius=nonMaxSuppRadius)
        self.T_w = []       # trajectory (camera -> world)
        self.points = []    # dicts: X, vis, patch, desc, x, y, saliency
        self.frame_id = 0

    def add_frame(self, I, Z, T_rel):
        o = self.o
        Tinv = np.linalg.inv(T_rel)
        self.T_w.append(self.T_w[-1] @ Tinv if self.T_w else Tinv)
        T_w = self.T_w[-1]
        T_c = np.eye(4); T_c[:3, :3] = T_w[:3, :3].T; T_c[:3, 3] = -T_w[:3, :3].T @ T_w[:3, 3]
        rows, cols = self.rows, self.cols
        B = max(o["maskBlockRadius"], max(2, o["patchRadius"]))
        max_rows, max_cols = rows - B - 1, cols - B - 1
        mask = np.ones((rows, cols), dtype=bool)
        for pt in self.points:
            if self.frame_id - pt["vis"][-1] <= o["maxFrameDistance"]:
                Xc = T_c[:3, :3] @ pt["X"] + T_c[:3, 3]
                p = self.K @ Xc
                u, v = p[0] / p[2], p[1] / p[2]
                r, c = int(np.floor(v + 0.5)) if v >= 0 else -int(np.floor(-v + 0.5)), int(np.floor(u + 0.5)) if u >= 0 else -int(np.floor(-u + 0.5))
                if B <= r < max_rows and B <= c <= max_cols:
                    if float(pt["patch"].score(Zncc(I, u, v))) > o["minScore"]:
                        pt["vis"].append(self.frame_id)
                        m = o["maskBlockRadius"]
                        mask[r - m:r + m + 1, c - m:c + m + 1] = False
        If = I.astype(f32)
        S = np.zeros((rows, cols), dtype=f32)
        S[1:-1, 1:-1] = np.abs(f32(0.5) * (If[1:-1, 2:] - If[1:-1, :-2])) + np.abs(f32(0.5) * (If[2:, 1:-1] - If[:-2, 1:-1]))
        n = o["nonMaxSuppRadius"]
        cand = np.zeros((rows, cols), dtype=bool)
        cand[B:max_rows, B:max_cols] = True
        cand &= (Z >= o["minValidDepth"]) & (Z <= o["maxValidDepth"])
        cand &= local_maxima(S, mask, n)
        new = []
        for y, x in zip(*np.nonzero(cand)):
            z = float(Z[y, x])
            v = np.array([float(x), float(y), 1.0])
            Xc = np.array([(z * self.Kinv[i, 0]) * v[0] + (z * self.Kinv[i, 1]) * v[1] + (z * self.Kinv[i, 2]) * v[2] for i in range(3)])
            X = T_w[:3, :3] @ Xc + T_w[:3, 3]
            new.append(dict(X=X, vis=[self.frame_id], patch=Zncc(I, float(x), float(y)), x=int(x), y=int(y),
                            saliency=float(S[y, x]), desc=None))
        if len(new) > o["maxNumPoints"]:
            new.sort(key=lambda p: -p["saliency"])
            self.dropped_max = new[o["maxNumPoints"]]["saliency"]
            new = new[:o["maxNumPoints"]]
        for p in new:
            p["desc"] = extract_patch(I, p["x"], p["y"], o["patchRadius"])
        self.points += new
        self.frame_id += 1
