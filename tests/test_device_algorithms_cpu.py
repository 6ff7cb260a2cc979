"""CPU emulations of the integer / bit-level tricks the CUDA kernels rely on, checked against the oracle restatements
(no GPU needed): the dp2a coefficient split of the K9 vertical pass (csrc/imgproc.cu crop_resize_up_kernel) and the
word-level 3x3 morphology of the line branch (csrc/refine.cu morph3x3_kernel)."""
import numpy as np


def test_dp2a_coefficient_split_is_exact():
    """k = kh * 2^11 + kl with kl = k & 2047, kh = k >> 11 (arithmetic): sum k * v == (sum kh * v << 11) + sum kl * v, both
    halves fit the signed 16-bit lanes of dp2a, for Pillow's whole coefficient range (negative lobes included)."""
    from oracle import resample
    rng = np.random.default_rng(0)
    for in_size in (1, 2, 3, 17, 45, 63, 64, 200, 383, 384):
        _, kk = resample.pil_coeffs(in_size, 384)
        k = np.asarray(kk, np.int64)
        kh, kl = k >> 11, k & 2047
        assert (k == kh * 2048 + kl).all() and kh.min() >= -32768 and kh.max() <= 32767 and kl.min() >= 0
        v = rng.integers(0, 256, k.shape, dtype=np.int64)
        full = (k * v).sum(1) + (1 << 21)
        split = ((kh * v).sum(1) << 11) + (kl * v).sum(1) + (1 << 21)
        assert (full == split).all()
        assert np.abs((kh * v).sum(1)).max() < 2 ** 31 and (kl * v).sum(1).max() < 2 ** 31
        # the clamp folded into the 512-entry table: (acc >> 22) + 128 stays inside it
        idx = (full >> 22) + 128
        assert idx.min() >= 0 and idx.max() < 512


def _to_words(mask):
    h, w = mask.shape
    wd = (w + 31) // 32
    pad = np.zeros((h, wd * 32), bool)
    pad[:, :w] = mask
    bits = pad.reshape(h, wd, 32)
    return (bits * (1 << np.arange(32, dtype=np.uint64))).sum(2).astype(np.uint32)


def _from_words(words, w):
    h, wd = words.shape
    bits = ((words[:, :, None].astype(np.uint64) >> np.arange(32, dtype=np.uint64)) & 1).astype(bool)
    return bits.reshape(h, wd * 32)[:, :w]


def _morph_words(words, w, erode):
    """morph3x3_kernel, word for word."""
    h, wd = words.shape
    full = np.uint32(0xFFFFFFFF)
    tail = np.uint32((1 << (w & 31)) - 1) if (w & 31) else full
    out = np.zeros_like(words)
    for y in range(h):
        for wx in range(wd):
            acc = full if erode else np.uint32(0)
            for dy in (-1, 0, 1):
                yy = y + dy
                if yy < 0 or yy >= h:
                    continue
                m = words[yy, wx]
                left = words[yy, wx - 1] if wx > 0 else (full if erode else np.uint32(0))
                right = words[yy, wx + 1] if wx + 1 < wd else (full if erode else np.uint32(0))
                if erode:
                    if wx == wd - 1:
                        m = m | ~tail
                    if wx + 1 == wd - 1:
                        right = right | ~tail
                l1 = np.uint32((int(m) << 1) & 0xFFFFFFFF) | (left >> np.uint32(31))
                r1 = (m >> np.uint32(1)) | np.uint32((int(right) << 31) & 0xFFFFFFFF)
                acc = (acc & m & l1 & r1) if erode else (acc | m | l1 | r1)
            if wx == wd - 1:
                acc = acc & tail
            out[y, wx] = acc
    return out


def test_word_level_closing_matches_restatement():
    from oracle import craft_post
    rng = np.random.default_rng(1)
    for (h, w), p in (((9, 31), 0.5), ((12, 32), 0.7), ((7, 33), 0.4), ((20, 70), 0.6), ((5, 64), 0.3), ((1, 100), 0.5)):
        m = rng.random((h, w)) > p
        words = _to_words(m)
        closed = _morph_words(_morph_words(words, w, erode=False), w, erode=True)
        assert np.array_equal(_from_words(closed, w), craft_post.close3x3_restated(m)), (h, w)


def _run_start(row_words, wx, b):
    """run_start() of csrc/ccl.cu: first pixel of the run containing bit b of word wx."""
    m = int(row_words[wx])
    inv = (~m) & ((1 << b) - 1)
    if inv:
        return wx * 32 + inv.bit_length()
    for j in range(wx - 1, -1, -1):
        iv = (~int(row_words[j])) & 0xFFFFFFFF
        if iv:
            return j * 32 + iv.bit_length()
    return 0


def _run_ccl(mask):
    """The device labelling (ccl.cu) replayed sequentially: run nodes on bit words, union by smallest index, roots ranked in
    raster order, labels per pixel."""
    h, w = mask.shape
    words = _to_words(mask)
    wd = words.shape[1]
    parent = {}
    for y in range(h):                                   # ccl_mask_kernel / ccl_init_runs_kernel: one node per run start
        for wx in range(wd):
            m = int(words[y, wx])
            prev = (int(words[y, wx - 1]) >> 31) if wx else 0
            starts = m & ~(((m << 1) | prev) & 0xFFFFFFFF)
            for b in range(32):
                if (starts >> b) & 1:
                    parent[y * w + wx * 32 + b] = y * w + wx * 32 + b

    def find(a):
        while parent[a] != a:
            a = parent[a]
        return a

    for y in range(1, h):                                # ccl_merge_kernel
        for wx in range(wd):
            cur, up = int(words[y, wx]), int(words[y - 1, wx])
            both = cur & up
            if not both:
                continue
            if (both & 1) and wx > 0 and (int(words[y, wx - 1]) & int(words[y - 1, wx - 1]) & 0x80000000):
                rest = (~both) & 0xFFFFFFFF
                both = (both & ~((rest & -rest) - 1)) if rest else 0
            while both:
                b = (both & -both).bit_length() - 1
                a, c = find(y * w + _run_start(words[y], wx, b)), find((y - 1) * w + _run_start(words[y - 1], wx, b))
                if a != c:
                    parent[max(a, c)] = min(a, c)
                ln = 0
                while b + ln < 32 and (both >> (b + ln)) & 1:
                    ln += 1
                both &= ~(((1 << ln) - 1) << b)
    roots = sorted(k for k in parent if find(k) == k)    # flatten + rowscan + assign: raster rank of the root pixels
    rank = {r: i + 1 for i, r in enumerate(roots)}
    labels = np.zeros((h, w), np.int32)
    for y in range(h):                                   # ccl_finalize_kernel
        for wx in range(wd):
            m = int(words[y, wx])
            for b in range(32):
                if (m >> b) & 1 and wx * 32 + b < w:
                    s0 = _run_start(words[y], wx, b)
                    labels[y, wx * 32 + b] = rank[find(y * w + s0)]
    return labels, len(roots) + 1


def test_run_based_labelling_matches_opencv():
    """The run / bit-plane union-find with raster-rank renumbering gives cv2.connectedComponents' labels (4-connectivity),
    including runs that cross word boundaries, full words and widths that are not multiples of 32."""
    import cv2
    rng = np.random.default_rng(2)
    for (h, w), p in (((12, 31), 0.5), ((16, 64), 0.35), ((10, 70), 0.25), ((24, 97), 0.45), ((6, 130), 0.1)):
        m = rng.random((h, w)) > p
        n_ref, ref = cv2.connectedComponents(m.astype(np.uint8), connectivity=4)
        lab, n = _run_ccl(m)
        assert n == n_ref and np.array_equal(lab, ref), (h, w)


def test_k1_source_row_walk_emits_every_row_once():
    """page_preprocess_kernel (csrc/imgproc.cu) walks the staged source rows of a 16-row output tile once and emits every
    output row whose LOWER source row is the current one.  Replayed here for many (source, target) heights — identity,
    down-scales up to the staging limit, the clamped bottom edge — every output row must come out exactly once, from the
    source-row pair cv2's INTER_LINEAR table prescribes."""
    from oracle import resample
    rng = np.random.default_rng(3)
    cases = [(330, 255), (3300, 2550), (200, 200), (17, 16), (64, 33), (1000, 435)]
    cases += [(int(s), int(max(1, s * f))) for s, f in zip(rng.integers(20, 900, 30), rng.uniform(0.44, 1.0, 30))]
    for sh, th in cases:
        ofs, _, _ = resample._cv_lin_coeffs(sh, th)
        for oy0 in range(0, th, 16):
            n_rows = min(oy0 + 16, th) - oy0
            ys0 = int(ofs[oy0])
            nsrc = min(int(ofs[oy0 + n_rows - 1]) + 1, sh - 1) - ys0 + 1
            assert nsrc <= 40, (sh, th)                       # K1_MAX_SRC_ROWS for factors the launcher accepts
            k, emitted = 0, []
            for r in range(nsrc):
                while k < n_rows:
                    o = int(ofs[oy0 + k])
                    if min(o + 1, sh - 1) - ys0 != r:
                        break
                    same = (o - ys0 == r)
                    emitted.append((oy0 + k, ys0 + (r if same else r - 1), ys0 + r))
                    k += 1
            want = [(oy, int(ofs[oy]), min(int(ofs[oy]) + 1, sh - 1)) for oy in range(oy0, oy0 + n_rows)]
            assert emitted == want, (sh, th, oy0)


def test_k9_pair_table_and_bands_reproduce_pillow_vertical_pass():
    """crop_resize_up_kernel (csrc/imgproc.cu): the vertical taps of every output row are re-packed into row PAIRS
    (p0, pairs, three (k_even, k_odd) slots) and the rows are processed in bands whose source pairs fit the staging area.
    Replayed for every crop height 1..384: the pair form must give Pillow's vertical pass exactly, and every band must stay
    inside its allocation."""
    from oracle import resample
    rng = np.random.default_rng(4)
    for h in list(range(1, 70)) + [95, 96, 97, 128, 200, 255, 256, 383, 384]:
        bounds, kk = resample.pil_coeffs(h, 384)
        assert kk.shape[1] == 5 and int(bounds[:, 1].max()) <= 5            # up-scale: at most five taps
        tmp = rng.integers(0, 256, (h, 7), dtype=np.int64)                   # a few columns of the u8 intermediate
        npair = (h + 1) // 2
        rows = np.zeros((2 * npair + 6, 7), np.int64)                        # pair storage; rows past the crop: arbitrary
        rows[:h] = tmp
        rows[h:] = rng.integers(0, 256, rows[h:].shape)
        table = []
        for yy in range(384):
            ymin, n = int(bounds[yy, 0]), int(bounds[yy, 1])
            odd = ymin & 1
            k6 = [0] * 6
            for x in range(n):
                k6[x + odd] = int(kk[yy, x])
            p0, pairs = ymin >> 1, (odd + n + 1) >> 1
            assert 1 <= pairs <= 3
            table.append((p0, pairs, k6))
            acc = 1 << 21
            for q in range(pairs):
                acc = acc + k6[2 * q] * rows[2 * (p0 + q)] + k6[2 * q + 1] * rows[2 * (p0 + q) + 1]
            want = (kk[yy, :n].astype(np.int64)[:, None] * tmp[ymin:ymin + n]).sum(0) + (1 << 21)
            assert (acc == want).all(), (h, yy)
        # band selection with the smallest staging area the kernel accepts (UP_MIN_PAIRS) and with a typical one
        for max_pairs in (12, 38):
            yb0 = 0
            while yb0 < 384:
                pf = table[yb0][0]
                yb1 = yb0 + 8
                while yb1 < 384 and table[yb1 + 7][0] + table[yb1 + 7][1] - pf <= max_pairs:
                    yb1 += 8
                used = max(t[0] + t[1] for t in table[yb0:yb1]) - pf
                assert 0 < used <= max_pairs, (h, yb0, yb1, used)
                assert all(t[0] >= pf for t in table[yb0:yb1])
                yb0 = yb1


def test_row_extremes_from_text_bit_runs():
    """box_extract_small_kernel (csrc/boxes.cu) fetches ONE label per run of consecutive `text > low_text` bits inside a
    component's bounding box (such a run lies inside one foreground run, hence inside one component).  Replayed against the
    direct definition min / max x of {label == k and text > low_text} per row."""
    import cv2
    from oracle import synth
    for seed in (1, 2, 3):
        text, link = synth.random_score_maps(seed, 90, 150, n_blobs=25)
        fg = (text > np.float32(0.3)) | (link > np.float32(0.45))
        tx = text > np.float32(0.3)
        n, labels, stats, _ = cv2.connectedComponentsWithStats(fg.astype(np.uint8), connectivity=4)
        words = _to_words(tx)
        for k in range(1, n):
            x0, y0, bw, bh = (int(stats[k, i]) for i in range(4))
            for r in range(bh):
                y = y0 + r
                mn, mx = 0x7fff, -1
                for wx in range(x0 >> 5, ((x0 + bw - 1) >> 5) + 1):
                    t = int(words[y, wx])
                    lo, hi = x0 - wx * 32, x0 + bw - 1 - wx * 32
                    if lo > 0:
                        t &= (0xFFFFFFFF << lo) & 0xFFFFFFFF
                    if hi < 31:
                        t &= 0xFFFFFFFF >> (31 - hi)
                    while t:
                        b = (t & -t).bit_length() - 1
                        ln = 0
                        while b + ln < 32 and (t >> (b + ln)) & 1:
                            ln += 1
                        if labels[y, wx * 32 + b] == k:
                            mn, mx = min(mn, wx * 32 + b), max(mx, wx * 32 + b + ln - 1)
                        t &= ~(((1 << ln) - 1) << b)
                xs = np.nonzero((labels[y, x0:x0 + bw] == k) & tx[y, x0:x0 + bw])[0]
                want = (x0 + int(xs.min()), x0 + int(xs.max())) if len(xs) else (0x7fff, -1)
                assert (mn, mx) == want, (seed, k, r)


def test_k9_v2_quads_digits_windows_reproduce_pillow():
    """Replay of crop_resize_up2_kernel (csrc/imgproc.cu) in numpy: at most FOUR taps on an up-scaled axis (exhaustive over
    in_size 1..384), the 16 + 8 bit coefficient split (kl = k & 0xFFFF for dp2a, kh = k >> 16 a signed byte for dp4a), shifted four-column
    horizontal window, row-quad words + byte-select windows, runs of rows with one window and the band walk over quads —
    equal to the oracle's Pillow restatement on random crops, including bands (small quad budget) and borders."""
    from oracle import resample
    # 1. four taps, digits in range
    for in_size in range(1, 385):
        b, kk = resample.pil_coeffs(in_size, 384)
        assert b[:, 1].max() <= 4
        k = np.asarray(kk, np.int64)
        assert (k[:, 4:] == 0).all()
        k2 = k >> 16
        assert k2.min() >= -128 and k2.max() <= 127
        assert (k == (k2 << 16) + (k & 0xFFFF)).all()
        # window of four source rows / columns never leaves the axis once shifted (in_size >= 4)
        if in_size >= 4:
            hx = np.minimum(b[:, 0], in_size - 4)
            assert (hx >= 0).all() and (b[:, 0] + b[:, 1] <= hx + 4).all()

    def replay(img, QB):
        h, w, _ = img.shape
        bv, kv = resample.pil_coeffs(h, 384)
        bh, kh = resample.pil_coeffs(w, 384)
        kv = np.asarray(kv, np.int64)[:, :4]
        kh = np.asarray(kh, np.int64)[:, :4]
        # horizontal pass through the shifted window
        hx = np.minimum(bh[:, 0], w - 4)
        sh = bh[:, 0] - hx
        hk = np.zeros((384, 4), np.int64)
        for x in range(384):
            for j in range(4):
                if 0 <= j - sh[x] < 4:
                    hk[x, j] = kh[x, j - sh[x]]
        nq_total = (h + 3) // 4
        ymin = bv[:, 0].astype(np.int64)
        first = np.ones(384, bool)
        first[1:] = ymin[1:] != ymin[:-1]
        run_y0 = list(np.nonzero(first)[0]) + [384]
        n_runs = len(run_y0) - 1
        out = np.zeros((384, 384, 3), np.uint8)
        klk, khk = kv & 0xFFFF, kv >> 16
        rb0 = 0
        while rb0 < n_runs:
            bq0 = int(ymin[run_y0[rb0]] >> 2)
            rb1 = rb0
            while rb1 < n_runs and (ymin[run_y0[rb1]] >> 2) <= bq0 + QB - 2:
                rb1 += 1
            assert rb1 > rb0
            nq = min(bq0 + QB, nq_total) - bq0
            tmp = np.full((QB + 1, 384 * 3, 4), 173, np.int64)          # garbage where nothing is computed
            for q in range(nq):
                for r in range(4):
                    row = min((bq0 + q) * 4 + r, h - 1)
                    src = img[row].astype(np.int64)
                    cols = hx[:, None] + np.arange(4)[None, :]
                    acc = (src[cols] * hk[:, :, None]).sum(1) + (1 << 21)
                    tmp[q, :, r] = np.clip(acc >> 22, 0, 255).reshape(-1)
            for r in range(rb0, rb1):
                y0, y1 = run_y0[r], run_y0[r + 1]
                q0, off = int(ymin[y0] >> 2) - bq0, int(ymin[y0] & 3)
                both = np.concatenate([tmp[q0], tmp[q0 + 1]], axis=1)   # 8 bytes: lo quad, hi quad
                win = both[:, off:off + 4]
                for yy in range(y0, y1):
                    dl = (win * klk[yy]).sum(1) + (1 << 21)
                    assert dl.max() < 2 ** 31
                    tot = ((win * khk[yy]).sum(1) << 16) + dl
                    idx = (tot >> 22) + 128
                    assert idx.min() >= 0 and idx.max() < 512
                    out[yy] = np.clip(tot >> 22, 0, 255).reshape(384, 3)
            rb0 = rb1
        return out

    rng = np.random.default_rng(5)
    for (h, w, QB) in [(63, 170, 21), (63, 170, 5), (5, 4, 13), (1, 9, 13), (384, 384, 13), (200, 33, 6), (97, 384, 13), (30, 7, 2)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = resample.pil_bicubic_resize_u8(img)
        assert np.array_equal(replay(img, QB), ref), (h, w, QB)


def test_k9_normalise_as_one_fma_is_exact():
    """crop_resize_up2_kernel replaces the (u / 255 - 0.5) / 0.5 table by fma(u, fl32(2/255), -1) rounded to the 16-bit
    element type: identical for all 256 pixel values in fp16 and bf16 (the fp32 values themselves differ)."""
    import torch
    u = np.arange(256, dtype=np.float32)
    ref32 = ((u / np.float32(255)) - np.float32(0.5)) / np.float32(0.5)
    a = np.float32(2.0 / 255.0)
    y = (u.astype(np.float64) * np.float64(a) - 1.0).astype(np.float32)      # exact product + one rounding = fma
    for dt in (torch.float16, torch.bfloat16):
        assert torch.equal(torch.from_numpy(y).to(dt), torch.from_numpy(ref32).to(dt))


# ------------------------------------------------------------------------------------------------ xattn_tc.cu layouts
def _sw128(byte_off):
    """128-byte swizzle of TMA / tcgen05 descriptors: bits [4,7) of a shared-memory offset XOR bits [7,10)."""
    return byte_off ^ (((byte_off >> 7) & 7) << 4)


def test_cross_attention_tile_serves_both_mmas_and_the_k9_window_gather():
    """dec_cross_tc_kernel (csrc/xattn_tc.cu) reads ONE shared-memory encoder tile twice: K-major for S^T = E_tile . q'^T and
    MN-major for ctx^T += E_tile^T . P^T.  Emulates the byte layout the TMA writes (boxes of [KT keys x 64 columns], 128-byte
    swizzle), the canonical K-major / MN-major descriptor walks (SBO = 1024; MN-major LBO = box stride), the P^T tile the
    softmax threads write (K-major, [16 x beams] rows x 64 keys) and the smem-free column -> q' row mapping, and checks that
    the two contractions computed through these address maps equal plain matrix products.  Also the PRMT selectors of the
    K9 horizontal pass (csrc/imgproc.cu): bytes c, 3 + c, 6 + c, 9 + c of a 12-byte window."""
    rng = np.random.default_rng(3)
    for E, KT, NB in ((768, 64, 1), (768, 32, 3), (1024, 32, 2), (128, 64, 2)):
        N, NCH, CHUNK = 16 * NB, E // 64, KT * 128
        enc = rng.standard_normal((KT, E)).astype(np.float32)
        q = rng.standard_normal((N, E)).astype(np.float32)
        # --- what the TMA leaves in shared memory (16-bit elements, modelled as a dict byte offset -> value)
        tile = {}
        for c in range(NCH):
            for r in range(KT):
                for e in range(64):
                    tile[c * CHUNK + _sw128(r * 128 + e * 2)] = enc[r, c * 64 + e]
        qs = {}
        for c in range(NCH):
            for r in range(N):
                for e in range(64):
                    qs[c * N * 128 + _sw128(r * 128 + e * 2)] = q[r, c * 64 + e]
        # --- S^T: A K-major (row m of the MMA = key, 8-row groups SBO = 1024 apart, k-step = +32 B), B = q' K-major
        S = np.zeros((KT, N), np.float32)
        for c in range(NCH):
            for k in range(4):
                for kk in range(16):
                    a = np.array([tile[c * CHUNK + _sw128((m // 8) * 1024 + (m % 8) * 128 + k * 32 + kk * 2)] for m in range(KT)])
                    b = np.array([qs[c * N * 128 + _sw128((n // 8) * 1024 + (n % 8) * 128 + k * 32 + kk * 2)] for n in range(N)])
                    S += np.outer(a, b)
        assert np.allclose(S, enc @ q.T, rtol=1e-4, atol=1e-3)
        # --- P^T as the softmax threads store it: thread t (key), column col -> row col of a K-major [N x 64 keys] tile
        P = rng.random((KT, N)).astype(np.float32)
        ps = {}
        for t in range(KT):
            for col in range(N):
                ps[col * 128 + (((t >> 3) ^ (col & 7)) << 4) + (t & 7) * 2] = P[t, col]
        # --- ctx^T: A MN-major (m = e column: 64-element atoms LBO = CHUNK apart; k = key: 8-row groups SBO = 1024, a k-step of
        #     16 keys = +2048 B), B = P^T K-major (k-step = +32 B)
        ctx = np.zeros((E, N), np.float32)
        for mt in range(E // 128):
            for k in range(KT // 16):
                for kk in range(16):
                    base = 2 * mt * CHUNK + k * 2048
                    a = np.array([tile[base + (m // 64) * CHUNK + _sw128((kk // 8) * 1024 + (kk % 8) * 128 + (m % 64) * 2)]
                                  for m in range(128)])
                    b = np.array([ps[_sw128((n // 8) * 1024 + (n % 8) * 128 + k * 32 + kk * 2)] for n in range(N)])
                    ctx[mt * 128:(mt + 1) * 128] += np.outer(a, b)
        assert np.allclose(ctx, enc.T @ P, rtol=1e-4, atol=1e-3)
    # --- work items: (crop, group) -> first q' row and the rows that are stored, heads = 16 and heads = 2
    for beam, NB, heads in ((5, 3, 16), (3, 3, 16), (3, 2, 16), (5, 3, 2), (8, 3, 16)):
        groups = -(-beam // NB)
        seen = []
        for crop in range(3):
            for grp in range(groups):
                b0 = grp * NB
                row0 = (crop * beam + b0) * heads
                row_end = (crop * beam + min(b0 + NB, beam)) * heads
                seen += [row0 + c for c in range(16 * NB) if row0 + c < row_end]
        assert seen == list(range(3 * beam * heads))             # every (crop, hypothesis, head) row exactly once, in order
    # --- K9 horizontal pass: PRMT selectors gather the four taps of a channel from three little-endian words
    def prmt(a, b, sel):
        by = [(a >> (8 * i)) & 255 for i in range(4)] + [(b >> (8 * i)) & 255 for i in range(4)]
        return sum(by[(sel >> (4 * i)) & 7] << (8 * i) for i in range(4))
    w = rng.integers(0, 256, 12).tolist()
    r0, r1, r2 = (sum(w[4 * j + i] << (8 * i) for i in range(4)) for j in range(3))
    for c, (s1, s2) in enumerate(((0x0630, 0x5210), (0x0741, 0x6210), (0x0052, 0x7410))):
        v = prmt(prmt(r0, r1, s1), r2, s2)
        assert [(v >> (8 * i)) & 255 for i in range(4)] == [w[c], w[3 + c], w[6 + c], w[9 + c]]
