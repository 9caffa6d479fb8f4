"""Run the hardware probes on a B200 and record which shared-memory layout hypotheses hold.

Usage (GPU box): python probes/run_probe.py  -> gpurun_out/probe.json + stdout table.
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = ctypes.CDLL(os.path.join(HERE, "libprobe.so"))


class UmmaArgs(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "img_bytes", "a_off", "a_lbo", "a_sbo", "a_layout", "a_base_mode",
        "b_off", "b_lbo", "b_sbo", "b_layout", "b_base_mode",
        "idesc", "nsteps", "a_step", "b_step", "N")]


LIB.probe_umma.argtypes = [ctypes.c_void_p, ctypes.POINTER(UmmaArgs), ctypes.c_void_p]
LIB.probe_umma.restype = ctypes.c_int
LIB.probe_tma.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                          ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p]
LIB.probe_tma.restype = ctypes.c_int
LIB.probe_tma_bw.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_int] * 8
LIB.probe_tma_bw.restype = ctypes.c_double

SWZ_NONE, SWZ_128, SWZ_64, SWZ_32 = 0, 2, 4, 6
ROWB = {SWZ_128: 128, SWZ_64: 64, SWZ_32: 32}


def f2bf(x):
    """float32 array (exactly representable small ints) -> bf16 bit patterns (uint16)."""
    return (np.asarray(x, np.float32).view(np.uint32) >> 16).astype(np.uint16)


def idesc(M, N, a_mn=0, b_mn=0):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24)


def swz_chunk(row_addr, c, layout):
    """physical 16B chunk index for logical chunk c of a row starting at byte address row_addr (absolute model)."""
    if layout == SWZ_128:
        return c ^ ((row_addr >> 7) & 7)
    if layout == SWZ_64:
        return c ^ ((row_addr >> 7) & 3)
    if layout == SWZ_32:
        return c ^ ((row_addr >> 7) & 1)
    return c


def put_rows(img, base, mat_u16, layout):
    """Store a [rows][rowbytes/2] bf16 matrix as consecutive rows of ROWB bytes with the absolute-address swizzle."""
    rb = ROWB[layout]
    rows, cols = mat_u16.shape
    assert cols * 2 == rb
    for r in range(rows):
        ra = base + r * rb
        for c in range(rb // 16):
            pc = swz_chunk(ra, c, layout)
            img[ra + pc * 16: ra + pc * 16 + 16] = mat_u16[r, c * 8:(c + 1) * 8].view(np.uint8)


def run_umma(img, **kw):
    a = UmmaArgs()
    a.img_bytes = img.size
    for k, v in kw.items():
        setattr(a, k, v)
    out = np.zeros((128, a.N), np.float32)
    rc = LIB.probe_umma(img.ctypes.data, ctypes.byref(a), out.ctypes.data)
    if rc != 0:
        raise RuntimeError("probe_umma rc=%d" % rc)
    return out


RNG = np.random.default_rng(0)
RESULTS = {}


def rnd(*shape):
    return RNG.integers(-3, 4, size=shape).astype(np.float32)


def record(name, ok, extra=None):
    RESULTS[name] = {"ok": bool(ok), **(extra or {})}
    print("%-64s %s %s" % (name, "PASS" if ok else "FAIL", extra or ""), flush=True)


# ----------------------------------------------------------------------------------------------
def test_kmajor_swizzled(layout, N, shift_rows, group_stride_rows, base_mode, tag):
    """A = 128-row view into a plane of swizzled rows: logical row m -> plane row shift + (m//8)*gs + m%8."""
    rb = ROWB[layout]
    kpr = rb // 2  # K elements per row
    plane_rows = shift_rows + 16 * group_stride_rows + 8
    plane = rnd(plane_rows, kpr)
    B = rnd(N, kpr)
    a_bytes = (plane_rows * rb + 1023) // 1024 * 1024
    img = np.zeros(a_bytes + N * rb + 1024, np.uint8)
    put_rows(img, 0, f2bf(plane), layout)
    put_rows(img, a_bytes, f2bf(B), layout)
    rows = np.array([shift_rows + (m // 8) * group_stride_rows + (m % 8) for m in range(128)])
    A = plane[rows]
    ref = A @ B.T
    out = run_umma(img, a_off=shift_rows * rb, a_lbo=16, a_sbo=group_stride_rows * rb, a_layout=layout,
                   a_base_mode=base_mode, b_off=a_bytes, b_lbo=16, b_sbo=8 * rb, b_layout=layout, b_base_mode=0,
                   idesc=idesc(128, N), nsteps=kpr // 16, a_step=32, b_step=32, N=N)
    ok = np.array_equal(out, ref)
    record(tag, ok, {"maxerr": float(np.abs(out - ref).max())})
    return ok


def test_kmajor_interleaved(N, shift_rows, group_stride_rows, swap, tag):
    """No-swizzle K-major: [kchunk][row][16B]; LBO = kchunk stride, SBO = 8-row-group stride (swap tests the reverse)."""
    K = 64
    plane_rows = shift_rows + 16 * group_stride_rows + 8
    plane = rnd(plane_rows, K)
    B = rnd(N, K)
    a_lbo = plane_rows * 16
    a_bytes = (K // 8) * a_lbo
    a_bytes_al = (a_bytes + 127) // 128 * 128
    b_lbo = N * 16
    img = np.zeros(a_bytes_al + (K // 8) * b_lbo + 256, np.uint8)
    pu = f2bf(plane)
    bu = f2bf(B)
    for kc in range(K // 8):
        for r in range(plane_rows):
            o = kc * a_lbo + r * 16
            img[o:o + 16] = pu[r, kc * 8:(kc + 1) * 8].view(np.uint8)
        for r in range(N):
            o = a_bytes_al + kc * b_lbo + r * 16
            img[o:o + 16] = bu[r, kc * 8:(kc + 1) * 8].view(np.uint8)
    rows = np.array([shift_rows + (m // 8) * group_stride_rows + (m % 8) for m in range(128)])
    ref = plane[rows] @ B.T
    a_sbo = group_stride_rows * 16
    b_sbo = 128
    kw = dict(a_off=shift_rows * 16, a_lbo=a_lbo, a_sbo=a_sbo, b_lbo=b_lbo, b_sbo=b_sbo)
    if swap:
        kw = dict(a_off=shift_rows * 16, a_lbo=a_sbo, a_sbo=a_lbo, b_lbo=b_sbo, b_sbo=b_lbo)
    out = run_umma(img, a_layout=SWZ_NONE, a_base_mode=0, b_off=a_bytes_al, b_layout=SWZ_NONE, b_base_mode=0,
                   idesc=idesc(128, N), nsteps=K // 16, a_step=2 * a_lbo, b_step=2 * b_lbo, N=N, **kw)
    ok = np.array_equal(out, ref)
    record(tag, ok, {"maxerr": float(np.abs(out - ref).max())})
    return ok


def test_mnmajor(layout, n_blocks, blk_stride_rows, shift_rows, kgroup_stride_rows, N, b_mn, tag):
    """wgrad-style operands. A is MN-major: smem rows are K (voxels), each row holds ROWB/2 channels (M).
    M = n_blocks * (ROWB/2): block j is the same plane shifted by j*blk_stride_rows rows (tap views).
    K = 16 per MMA = two 8-row groups kgroup_stride_rows apart; 2 MMAs (K=32) advance by 2 groups.
    B: K-major (b_mn=0, [N][K] rows swizzled) or MN-major (b_mn=1, rows are K, N channels per row)."""
    rb = ROWB[layout]
    mpr = rb // 2
    M = n_blocks * mpr
    assert M in (64, 128)
    Ktot = 32
    plane_rows = shift_rows + (n_blocks - 1) * blk_stride_rows + 4 * kgroup_stride_rows + 8
    plane = rnd(plane_rows, mpr)
    a_bytes = (plane_rows * rb + 1023) // 1024 * 1024
    img = np.zeros(a_bytes + 8192, np.uint8)
    put_rows(img, 0, f2bf(plane), layout)

    def krow(k):
        return (k // 8) * kgroup_stride_rows + k % 8

    A = np.zeros((M, Ktot), np.float32)
    for j in range(n_blocks):
        for k in range(Ktot):
            A[j * mpr:(j + 1) * mpr, k] = plane[shift_rows + j * blk_stride_rows + krow(k)]
    if b_mn:
        # B rows are K (contiguous groups of 8 rows, 4 groups), N channels per row -> layout by N*2 bytes
        blayout = {256: None, 128: SWZ_128, 64: SWZ_64, 32: SWZ_32}[N * 2]
        Bk = rnd(Ktot, N)
        put_rows(img, a_bytes, f2bf(Bk), blayout)
        Bmat = Bk.T
        bkw = dict(b_off=a_bytes, b_lbo=16, b_sbo=8 * N * 2, b_layout=blayout, b_step=16 * N * 2)
    else:
        Bmat = rnd(N, Ktot)  # [N][K], K-major rows of 64 B
        put_rows(img, a_bytes, f2bf(Bmat), SWZ_64)
        bkw = dict(b_off=a_bytes, b_lbo=16, b_sbo=8 * 64, b_layout=SWZ_64, b_step=32)
    ref = A @ Bmat.T
    out = run_umma(img, a_off=shift_rows * rb, a_lbo=blk_stride_rows * rb, a_sbo=kgroup_stride_rows * rb,
                   a_layout=layout, a_base_mode=0, b_base_mode=0, idesc=idesc(M, N, 1, b_mn), nsteps=2,
                   a_step=2 * kgroup_stride_rows * rb, N=N, **bkw)
    out = out[:M] if M == 128 else out[:64]
    ok = np.array_equal(out, ref[:out.shape[0]])
    record(tag, ok, {"maxerr": float(np.abs(out - ref[:out.shape[0]]).max())})
    return ok


def test_tma(C, layout_tma, tag, interleaved=False):
    """Box load with a -1 halo: what lands in smem?"""
    D, H, W = 3, 20, 12
    vol = np.arange(D * H * W * C, dtype=np.float32).reshape(D, H, W, C) % 251 + 1
    data = f2bf(vol)
    BW, BH = 10, 18
    coords = np.array([0, -1, -1, 1, 0], np.int32)
    if not interleaved:
        dims = np.array([C, W, H, D, 1], np.uint64)
        strides = np.array([C * 2, W * C * 2, H * W * C * 2, D * H * W * C * 2], np.uint64)
        box = np.array([C, BW, BH, 1, 1], np.uint32)
    else:
        dims = np.array([8, W, H, D, C // 8], np.uint64)
        strides = np.array([C * 2, W * C * 2, H * W * C * 2, 16], np.uint64)
        box = np.array([8, BW, BH, 1, C // 8], np.uint32)
    nbytes = BW * BH * C * 2
    out = np.zeros(nbytes, np.uint8)
    rc = LIB.probe_tma(data.ctypes.data, data.nbytes, dims.ctypes.data, strides.ctypes.data, box.ctypes.data,
                       layout_tma, coords.ctypes.data, nbytes, out.ctypes.data)
    if rc != 0:
        record(tag, False, {"rc": rc})
        return False
    got = out.view(np.uint16)
    exp = np.zeros(nbytes // 2, np.uint16)
    umma_layout = {0: SWZ_NONE, 1: SWZ_32, 2: SWZ_64, 3: SWZ_128}[layout_tma]
    for bh in range(BH):
        for bw in range(BW):
            h, w = bh - 1, bw - 1
            inb = 0 <= h < H and 0 <= w < W
            vals = data[1, h, w] if inb else np.zeros(C, np.uint16)
            if not interleaved:
                row = bh * BW + bw
                ra = row * C * 2
                for c in range(C // 8):
                    pc = swz_chunk(ra, c, umma_layout)
                    exp[(ra + pc * 16) // 2:(ra + pc * 16) // 2 + 8] = vals[c * 8:(c + 1) * 8]
            else:
                for cg in range(C // 8):
                    o = (cg * BH * BW + bh * BW + bw) * 8
                    exp[o:o + 8] = vals[cg * 8:(cg + 1) * 8]
    ok = np.array_equal(got, exp)
    record(tag, ok, {"mismatch": int((got != exp).sum())})
    return ok


def test_tma_bw(C, interleaved, swz, tag):
    D = H = W = 128
    N = 2
    if not interleaved:
        dims = np.array([C, W, H, D * N, 1], np.uint64)
        strides = np.array([C * 2, W * C * 2, H * W * C * 2, N * D * H * W * C * 2], np.uint64)
        box = np.array([C, 10, 18, 1, 1], np.uint32)
    else:
        dims = np.array([8, W, H, D * N, C // 8], np.uint64)
        strides = np.array([C * 2, W * C * 2, H * W * C * 2, 16], np.uint64)
        box = np.array([8, 10, 18, 1, C // 8], np.uint32)
    # bw kernel uses coords (0, i1*s1-1, i2*s2-1, i3*s3, 0)
    gbs = LIB.probe_tma_bw(N * D * H * W * C * 2, dims.ctypes.data, strides.ctypes.data, box.ctypes.data, swz, 16, 8, 256, 8, 16, 1,
                           148 * 2, 512)
    record(tag, gbs > 0, {"GBps": round(float(gbs), 1)})


def build_tests():
    T = []
    add = lambda f, *a, **k: T.append((f, a, k))
    for N in (32, 64, 128, 256):
        add(test_kmajor_swizzled, SWZ_128, N, 0, 8, 0, "kmajor_sw128_canonical_N%d" % N)
    add(test_kmajor_swizzled, SWZ_64, 32, 0, 8, 0, "kmajor_sw64_canonical_N32")
    add(test_kmajor_swizzled, SWZ_32, 32, 0, 8, 0, "kmajor_sw32_canonical_N32")
    for lay, nm in ((SWZ_128, "sw128"), (SWZ_64, "sw64"), (SWZ_32, "sw32")):
        add(test_kmajor_swizzled, lay, 32, 0, 10, 0, "kmajor_%s_gs10_shift0" % nm)
        for s_ in (1, 2, 3, 5, 8, 11, 21):
            add(test_kmajor_swizzled, lay, 32, s_, 10, 0, "kmajor_%s_gs10_shift%d_base0" % (nm, s_))
        for s_ in (1, 3, 11):
            add(test_kmajor_swizzled, lay, 32, s_, 10, 1, "kmajor_%s_gs10_shift%d_basecomputed" % (nm, s_))
        for s_ in (1, 3):
            add(test_kmajor_swizzled, lay, 32, s_, 8, 0, "kmajor_%s_gs8_shift%d_base0" % (nm, s_))
            add(test_kmajor_swizzled, lay, 32, s_, 8, 1, "kmajor_%s_gs8_shift%d_basecomputed" % (nm, s_))
    add(test_kmajor_interleaved, 32, 0, 8, False, "kmajor_none_canonical_lbo=kchunk")
    add(test_kmajor_interleaved, 32, 0, 8, True, "kmajor_none_canonical_lbo=group(swapped)")
    for s_ in (1, 3, 11):
        add(test_kmajor_interleaved, 32, s_, 10, False, "kmajor_none_gs10_shift%d" % s_)
    add(test_mnmajor, SWZ_128, 2, 64, 0, 8, 32, 0, "mnmajor_sw128_blocks_far_canonical_Bk")
    add(test_mnmajor, SWZ_128, 2, 1, 0, 8, 32, 0, "mnmajor_sw128_blocks_1row_apart_Bk")
    add(test_mnmajor, SWZ_128, 2, 1, 3, 10, 32, 0, "mnmajor_sw128_blocks_1row_shift3_gs10_Bk")
    add(test_mnmajor, SWZ_128, 2, 10, 3, 10, 32, 0, "mnmajor_sw128_blocks_10row_shift3_gs10_Bk")
    add(test_mnmajor, SWZ_64, 4, 64, 0, 8, 32, 0, "mnmajor_sw64_4blocks_far_canonical_Bk")
    add(test_mnmajor, SWZ_64, 4, 1, 0, 8, 32, 0, "mnmajor_sw64_4blocks_1row_apart_Bk")
    add(test_mnmajor, SWZ_64, 4, 1, 3, 10, 32, 0, "mnmajor_sw64_4blocks_1row_shift3_gs10_Bk")
    add(test_mnmajor, SWZ_128, 2, 1, 3, 10, 32, 1, "mnmajor_sw128_A_and_B_mn_N32")
    add(test_mnmajor, SWZ_128, 2, 1, 3, 10, 64, 1, "mnmajor_sw128_A_and_B_mn_N64")
    add(test_mnmajor, SWZ_64, 4, 1, 3, 10, 32, 1, "mnmajor_sw64_A_and_B_mn_N32")
    add(test_tma, 64, 3, "tma_box_sw128_C64")
    add(test_tma, 32, 2, "tma_box_sw64_C32")
    add(test_tma, 16, 1, "tma_box_sw32_C16")
    add(test_tma, 32, 0, "tma_box_interleaved_C32", interleaved=True)
    add(test_tma_bw, 64, False, 3, "tma_bw_sw128_C64")
    add(test_tma_bw, 32, False, 2, "tma_bw_sw64_C32")
    add(test_tma_bw, 32, True, 0, "tma_bw_interleaved_C32")
    return T


def child(start, path):
    tests = build_tests()
    for i in range(start, len(tests)):
        f, a, k = tests[i]
        with open(path, "a") as fh:
            fh.write(json.dumps({"begin": i, "name": a[-1] if isinstance(a[-1], str) else str(a)}) + "\n")
        before = set(RESULTS)
        f(*a, **k)
        new = {n: RESULTS[n] for n in RESULTS if n not in before}
        with open(path, "a") as fh:
            fh.write(json.dumps({"end": i, "results": new}) + "\n")


def main():
    import subprocess
    os.makedirs("gpurun_out", exist_ok=True)
    path = "gpurun_out/probe.jsonl"
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), path)
        return
    open(path, "w").close()
    n = len(build_tests())
    start = 0
    merged = {}
    while start < n:
        rc = subprocess.call([sys.executable, os.path.abspath(__file__), "--child", str(start)], timeout=600)
        last_begin, names = -1, {}
        for line in open(path):
            r = json.loads(line)
            if "begin" in r:
                last_begin = r["begin"]; names[r["begin"]] = r["name"]
            else:
                merged.update(r["results"]); last_begin = -1 if r["end"] == last_begin else last_begin
        if rc == 0 and last_begin == -1:
            break
        crashed = last_begin if last_begin >= 0 else start
        merged[names.get(crashed, "test%d" % crashed)] = {"ok": False, "crash": rc}
        print("CRASH in test", crashed, names.get(crashed), "rc", rc, flush=True)
        start = crashed + 1
    with open("gpurun_out/probe.json", "w") as f:
        json.dump(merged, f, indent=1)
    print("done", sum(r["ok"] for r in merged.values()), "/", len(merged))


if __name__ == "__main__":
    main()
