"""CPU: the C-ABI library builds, loads, and exports every symbol include/b200seg.h declares with the argument
lists the ctypes binding uses; host-side logic that needs no GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_decls():
    hdr = open(os.path.join(ROOT, "include", "b200seg.h")).read()
    return re.findall(r"\n(?:int|int64_t|size_t|const char\*)\s+(b200seg_\w+)\s*\(([^;]*?)\)\s*;", hdr, re.S)


def _code(arg):
    a = arg.strip()
    if a in ("void", ""):
        return ""
    if "b200seg_conv_geom" in a:
        return "g"
    if "*" in a:
        return "p"
    return {"int": "i", "int64_t": "l", "float": "f", "double": "d", "size_t": "z", "int32_t": "i", "unsigned long long": "Q"}[a.rsplit(" ", 1)[0].strip()]


def test_build_and_exports_match_header():
    import __graft_entry__ as ge
    ge.build()
    from b200seg import _lib
    lib = _lib.load()
    decls = _header_decls()
    assert len(decls) >= 36
    for name, args in decls:
        assert hasattr(lib, name), "missing export " + name
        codes = "".join(_code(a) for a in args.replace("\n", " ").split(","))
        bound = _lib.SIGNATURES.get(name, _lib.SIZE_FUNCS.get(name, _lib.INT64_FUNCS.get(name, "" if name in _lib.STRING_FUNCS else None)))
        assert bound == codes, (name, bound, codes)
    assert set(_lib.exported_symbols()) == {n for n, _ in decls}
    assert b"sm_100a" in lib.b200seg_version()


def test_argument_validation_without_gpu():
    """Validation happens before any launch, so bad arguments are reported on a CPU-only box too."""
    from b200seg import _lib
    lib = _lib.load()
    g = _lib.ConvGeom(1, 8, 8, 8, 16, 8, 8, 8, 16, 3, 1, 1, 1)
    g.od = 5  # inconsistent with the geometry
    rc = lib.b200seg_conv3d_fprop(ctypes.byref(g), 1, 16, 1, None, 1, 16, None, None, 0, None)
    assert rc == -1 and b"output extents" in lib.b200seg_last_error()
    with pytest.raises(_lib.B200SegError):
        _lib.call("b200seg_seg_counts", None, None, 10, None, None)


def test_missing_extension_fails_loudly(monkeypatch):
    from b200seg import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb200seg.so")
    with pytest.raises(_lib.B200SegError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_model_surface_matches_reference_keys(golden):
    from b200seg.models.three_d.unet3d import UNet3D
    g = golden("unet_f4_s32_b2")
    ref = {k[4:]: g[k].shape for k in g.files if k.startswith("sd0.")}
    net = UNet3D(in_channels=1, out_channels=2, init_features=4)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == ref
    assert sum(p.numel() for p in UNet3D(1, 2, 32).parameters()) == 22581250
    with pytest.raises(ValueError):
        net(torch.zeros(1, 1, 8, 8))


def test_grid_sampler_matches_oracle():
    from b200seg.inference import GridSampler
    from oracle import window
    import numpy as np
    for shape, patch, ov in [((512, 512, 256), (128,) * 3, (64,) * 3), ((512, 512, 256), (128,) * 3, (4, 4, 36)),
                             ((40, 36, 50), (16,) * 3, (4, 4, 6)), ((16, 16, 16), (16,) * 3, (0, 0, 0))]:
        s = GridSampler(shape, patch, ov)
        assert np.array_equal(s.locations.numpy(), window.grid_locations(shape, patch, ov))
    with pytest.raises(ValueError):
        GridSampler((8, 8, 8), (16, 16, 16), (0, 0, 0))
    with pytest.raises(ValueError):
        GridSampler((32, 32, 32), (16, 16, 16), (3, 4, 4))


def test_build_outputs_are_renamed_into_place_and_builds_are_serialised(tmp_path):
    """The ranks of one torchrun launch all call build(): outputs appear atomically and only one build runs at a time
    (a rank once loaded a half-written library while seven others were linking it)."""
    import multiprocessing as mp
    import time
    import __graft_entry__ as ge
    target = tmp_path / "out.bin"
    src = tmp_path / "in.bin"
    src.write_bytes(b"x" * 1000)
    ge._run_atomic(["cp", str(src), str(target)], str(target))
    assert target.read_bytes() == src.read_bytes() and not list(tmp_path.glob("out.bin.tmp.*"))
    with pytest.raises(Exception):
        ge._run_atomic(["cp", str(tmp_path / "missing"), str(target)], str(target))
    assert target.read_bytes() == src.read_bytes() and not list(tmp_path.glob("out.bin.tmp.*"))   # old file intact
    # two processes inside _BuildLock never overlap
    log = tmp_path / "log.txt"

    def worker(path):
        import __graft_entry__ as g
        with g._BuildLock():
            with open(path, "a") as f:
                f.write("enter\n")
            time.sleep(0.3)
            with open(path, "a") as f:
                f.write("exit\n")

    ctx = mp.get_context("fork")
    procs = [ctx.Process(target=worker, args=(str(log),)) for _ in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert log.read_text().split() == ["enter", "exit"] * 3
