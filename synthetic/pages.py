"""Deterministic synthetic inputs shared by the oracle, the tests and bench.py.

Nothing here is reference code: the reference ships no pages, weights or golden vectors for this path."""
import numpy as np
import cv2


INK_BGR = (160, 0, 0)   # dark-blue ink: separable from both the white page and the black (zero) canvas padding


def synth_page(page_index, height=3300, width=2550, scale=1.2, line_pitch=70, gap=40, margin=150,
               thickness=2, color=INK_BGR):
    """White BGR page with dark-blue Hershey-simplex words, seeded by 1000+page_index (≈500 words on letter)."""
    rng = np.random.default_rng(1000 + int(page_index))
    img = np.full((height, width, 3), 255, np.uint8)
    y = margin + 40
    words = 0
    while y < height - margin:
        x = margin
        while True:
            n = int(rng.integers(3, 10))
            word = "".join(chr(ord("A") + int(c)) for c in rng.integers(0, 26, n))
            (tw, th), _ = cv2.getTextSize(word, cv2.FONT_HERSHEY_SIMPLEX, scale, thickness)
            if x + tw > width - margin:
                break
            cv2.putText(img, word, (x, y), cv2.FONT_HERSHEY_SIMPLEX, scale, color, thickness, cv2.LINE_AA)
            x += tw + gap
            words += 1
        y += line_pitch
    return img, words


def dense_page(page_index, height=3300, width=2550):
    """Config-4 style high-density form page (~800 words)."""
    return synth_page(page_index, height, width, scale=0.9, line_pitch=52, gap=30, thickness=2)


def score_maps_from_page(page_bgr, hm_h, hm_w):
    """'Injected oracle maps' for stage-level post-processing runs: blurred glyph masks at heat-map resolution
    (text = clip(3*blur(1-gray)), link = clip(1.2*blur(text, wide))) giving realistic component densities."""
    gray = cv2.cvtColor(page_bgr, cv2.COLOR_BGR2GRAY).astype(np.float32) / 255.0
    small = cv2.resize(1.0 - gray, (hm_w, hm_h), interpolation=cv2.INTER_AREA)
    text = np.clip(3.0 * cv2.GaussianBlur(small, (9, 9), 3), 0, 1).astype(np.float32)
    link = np.clip(1.2 * cv2.GaussianBlur(text, (15, 3), 5), 0, 1).astype(np.float32)
    return text, link


def random_score_maps(seed, h, w, n_blobs=60, rotated=True):
    """Random rotated-ellipse blobs with smooth fall-off: exercises rotated min-area rects, merges through link
    regions, tiny components (area<10) and border-touching components."""
    rng = np.random.default_rng(seed)
    text = np.zeros((h, w), np.float32)
    link = np.zeros((h, w), np.float32)
    for _ in range(n_blobs):
        cx, cy = rng.integers(0, w), rng.integers(0, h)
        ax, ay = int(rng.integers(2, max(3, w // 8))), int(rng.integers(1, max(2, h // 16)))
        ang = float(rng.uniform(0, 180)) if rotated else 0.0
        peak = float(rng.uniform(0.35, 1.0))
        m = np.zeros((h, w), np.float32)
        cv2.ellipse(m, (int(cx), int(cy)), (ax, ay), ang, 0, 360, peak, -1)
        text = np.maximum(text, m)
        if rng.random() < 0.5:
            m2 = np.zeros((h, w), np.float32)
            cv2.ellipse(m2, (int(cx + ax), int(cy)), (max(1, ax // 2), max(1, ay // 2)), ang, 0, 360,
                        float(rng.uniform(0.3, 0.9)), -1)
            link = np.maximum(link, m2)
    for _ in range(n_blobs // 4):   # specks below the area filter
        x, y = rng.integers(0, w - 2), rng.integers(0, h - 2)
        text[y:y + int(rng.integers(1, 3)), x:x + int(rng.integers(1, 4))] = float(rng.uniform(0.5, 1.0))
    text = cv2.GaussianBlur(text, (5, 5), 1.0)
    link = cv2.GaussianBlur(link, (5, 5), 1.0)
    noise = rng.normal(0, 0.01, (h, w)).astype(np.float32)
    return (text + noise).astype(np.float32), (link - noise).astype(np.float32)
