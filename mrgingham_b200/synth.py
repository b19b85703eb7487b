"""Deterministic synthetic frames for tests and bench.py (NumPy only, no OpenCV, no GPU).

Board geometry follows the reference's board generator (generate-chessboard-fig.py:100-134):
an (N+3)x(N+3) grid of unit cells whose outer ring of squares is double-size, so the board
has N+1 squares per side and exactly NxN interior X-corners. Rendering follows SURVEY.md
section 8(d): white 230 / black 30 on background 128, board side ~0.8*min(W,H), in-plane
rotation U(-0.3,0.3) rad, each quad corner scaled by 1+U(-0.12,0.12), supersampled, additive
Gaussian noise (sigma 2), then the 3x3 box blur the reference CLI applies by default
(mrgingham-from-image.cc:106-111).
"""
import numpy as np

WHITE, BLACK, BACKGROUND = 230, 30, 128


def _cell_table(n):
    """colour of unit cell (cy, cx) of the (n+3)x(n+3) board; index n+3 on either axis = outside"""
    m = n + 3
    t = np.full((m + 1, m + 1), BACKGROUND, dtype=np.float32)
    t[:m, :m] = WHITE
    for cy in range(m):
        for cx in range(m):
            edge_y = cy < 2 or cy > n
            edge_x = cx < 2 or cx > n
            if edge_y and edge_x:
                black = False
            elif edge_y:
                black = (cx % 2 == 0) and 2 <= cx < n + 1
            elif edge_x:
                black = (cy % 2 == 0) and 2 <= cy < n + 1
            else:
                black = (cx + cy) % 2 == 1
            if black:
                t[cy, cx] = BLACK
    return t


def _homography(src, dst):
    """3x3 H with H @ [src,1] ~ [dst,1] for four point pairs"""
    A, b = [], []
    for (x, y), (u, v) in zip(src, dst):
        A.append([x, y, 1, 0, 0, 0, -u * x, -u * y]); b.append(u)
        A.append([0, 0, 0, x, y, 1, -v * x, -v * y]); b.append(v)
    h = np.linalg.solve(np.asarray(A, dtype=np.float64), np.asarray(b, dtype=np.float64))
    return np.append(h, 1.0).reshape(3, 3)


def box_blur3(img):
    """3x3 box blur, rounded to nearest, edges replicated. uint8 in, uint8 out."""
    p = np.pad(img.astype(np.uint16), 1, mode="edge")
    s = np.zeros(img.shape, dtype=np.uint16)
    for dy in range(3):
        for dx in range(3):
            s += p[dy:dy + img.shape[0], dx:dx + img.shape[1]]
    return ((s * 2 + 9) // 18).astype(np.uint8)


def board_frame(w, h, n=10, seed=0, supersample=2, noise_sigma=2.0, blur=True):
    """One synthetic grayscale frame (h, w) uint8 holding an n x n-corner chessboard."""
    rng = np.random.default_rng(seed)
    m = n + 3
    side = 0.8 * min(w, h)
    theta = rng.uniform(-0.3, 0.3)
    c, s = np.cos(theta), np.sin(theta)
    centre = np.array([w / 2.0, h / 2.0])
    quad = []
    for ux, uy in ((-0.5, -0.5), (0.5, -0.5), (0.5, 0.5), (-0.5, 0.5)):
        k = 1.0 + rng.uniform(-0.12, 0.12)
        px, py = ux * side * k, uy * side * k
        quad.append((centre[0] + c * px - s * py, centre[1] + s * px + c * py))
    board = [(0.0, 0.0), (float(m), 0.0), (float(m), float(m)), (0.0, float(m))]
    H = _homography(quad, board).astype(np.float32)

    table = _cell_table(n)
    ss = supersample
    xs = ((np.arange(w * ss, dtype=np.float32) + 0.5) / ss - 0.5)[None, :]
    acc = np.empty((h, w), dtype=np.float32)
    rows_per_chunk = max(1, (1 << 22) // (w * ss * ss))
    for y0 in range(0, h, rows_per_chunk):
        y1 = min(h, y0 + rows_per_chunk)
        ys = ((np.arange(y0 * ss, y1 * ss, dtype=np.float32) + 0.5) / ss - 0.5)[:, None]
        den = H[2, 0] * xs + (H[2, 1] * ys + H[2, 2])
        bu = (H[0, 0] * xs + (H[0, 1] * ys + H[0, 2])) / den
        bv = (H[1, 0] * xs + (H[1, 1] * ys + H[1, 2])) / den
        cu = np.floor(bu).astype(np.int32)
        cv = np.floor(bv).astype(np.int32)
        outside = (cu < 0) | (cu >= m) | (cv < 0) | (cv >= m)
        cu[outside] = m
        cv[outside] = m
        val = table[cv, cu]
        acc[y0:y1] = val.reshape(y1 - y0, ss, w, ss).mean(axis=(1, 3))
    if noise_sigma > 0:
        acc += rng.normal(0.0, noise_sigma, size=acc.shape).astype(np.float32)
    img = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    if blur:
        img = box_blur3(img)
    return np.ascontiguousarray(img)


def noise_frame(w, h, seed=0):
    """uniform random bytes: ~3 % of pixels end up with response > 15"""
    return np.random.default_rng(seed).integers(0, 256, size=(h, w), dtype=np.uint8)


def blurred_noise_frame(w, h, seed=0, passes=2):
    img = noise_frame(w, h, seed)
    for _ in range(passes):
        img = box_blur3(img)
    f = img.astype(np.float32)
    f = (f - f.min()) * (255.0 / max(1.0, float(f.max() - f.min())))
    return np.rint(f).astype(np.uint8)


def checker_frame(w, h, period=8, seed=0, noise_sigma=1.0, blur=True):
    """axis-aligned dense checkerboard: thousands of X-corners, ~18 % of pixels > 15"""
    yy, xx = np.mgrid[0:h, 0:w]
    f = np.where(((xx // period) + (yy // period)) % 2 == 0, float(WHITE), float(BLACK)).astype(np.float32)
    if noise_sigma > 0:
        f += np.random.default_rng(seed).normal(0.0, noise_sigma, size=f.shape).astype(np.float32)
    img = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return box_blur3(img) if blur else img


def blob_frame(w, h, seed=0, nblobs=None):
    """random bright/dark discs on grey: irregular connected components of every size"""
    rng = np.random.default_rng(seed)
    f = np.full((h, w), float(BACKGROUND), dtype=np.float32)
    yy, xx = np.mgrid[0:h, 0:w]
    if nblobs is None:
        nblobs = max(8, (w * h) // 4000)
    for _ in range(nblobs):
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        r = rng.uniform(3, 18)
        v = float(rng.choice([BLACK, WHITE]))
        x0, x1 = max(0, int(cx - r - 1)), min(w, int(cx + r + 2))
        y0, y1 = max(0, int(cy - r - 1)), min(h, int(cy + r + 2))
        sub = (xx[y0:y1, x0:x1] - cx) ** 2 + (yy[y0:y1, x0:x1] - cy) ** 2 <= r * r
        f[y0:y1, x0:x1][sub] = v
    f += rng.normal(0.0, 2.0, size=f.shape).astype(np.float32)
    return box_blur3(np.clip(np.rint(f), 0, 255).astype(np.uint8))


def circle_grid_frame(w, h, n=10, seed=0, noise_sigma=2.0, blur=True):
    """n x n grid of dark discs on a light board (the calibration target mrgingham's blob mode is
    for: find_blobs.cc:22 'black-on-white dots'), mild rotation, noise, optional 3x3 box blur."""
    rng = np.random.default_rng(seed)
    side = 0.8 * min(w, h)
    theta = rng.uniform(-0.3, 0.3)
    c, s_ = np.cos(theta), np.sin(theta)
    pitch = side / (n + 1)
    r = pitch * rng.uniform(0.22, 0.32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    u = c * (xx - w / 2.0) + s_ * (yy - h / 2.0)
    v = -s_ * (xx - w / 2.0) + c * (yy - h / 2.0)
    f = np.full((h, w), float(BACKGROUND), dtype=np.float32)
    f[(np.abs(u) <= side / 2) & (np.abs(v) <= side / 2)] = float(WHITE)
    gu = (u + side / 2) / pitch
    gv = (v + side / 2) / pitch
    iu, iv = np.rint(gu), np.rint(gv)
    inside = (iu >= 1) & (iu <= n) & (iv >= 1) & (iv <= n)
    d2 = ((gu - iu) ** 2 + (gv - iv) ** 2) * pitch * pitch
    edge = np.clip(r + 0.5 - np.sqrt(d2), 0.0, 1.0)                     # one-pixel anti-aliased rim
    f = np.where(inside, f * (1 - edge) + float(BLACK) * edge, f)
    if noise_sigma > 0:
        f = f + rng.normal(0.0, noise_sigma, size=f.shape).astype(np.float32)
    img = np.clip(np.rint(f), 0, 255).astype(np.uint8)
    return box_blur3(img) if blur else img


def camera_frame(w, h, n=10, seed=0, noise_sigma=5.0):
    """A board the way a camera sees it: the synthetic board on a textured (blurred-noise) background, radial
    vignetting down to ~55 % in the corners, the 3x3 box blur, and THEN sensor noise of `noise_sigma` grey levels
    (so the noise reaches the detector unblurred, as shot noise does when the CLI runs with --blur 0)."""
    rng = np.random.default_rng(seed)
    board = board_frame(w, h, n, seed=seed, noise_sigma=0.0, blur=False).astype(np.float32)
    tex = blurred_noise_frame(w, h, seed=seed + 7919, passes=2).astype(np.float32)
    f = np.where(board == float(BACKGROUND), 0.5 * float(BACKGROUND) + 0.5 * tex, board)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    r2 = ((xx - w / 2.0) / (w / 2.0)) ** 2 + ((yy - h / 2.0) / (h / 2.0)) ** 2
    f = f * (1.0 - 0.225 * r2)
    img = box_blur3(np.clip(np.rint(f), 0, 255).astype(np.uint8)).astype(np.float32)
    img += rng.normal(0.0, noise_sigma, size=img.shape).astype(np.float32)
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))
