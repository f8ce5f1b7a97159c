"""The C-ABI library: builds for sm_100a, loads, and exports every symbol include/dcsb200.h
declares.  No compute calls here (no GPU in the build container)."""
import ctypes
import os
import re
import pytest
from conftest import ROOT


def test_library_exports_every_declared_symbol(built):
    from dcsexplorer_b200 import _capi
    L = _capi.lib()
    hdr = open(os.path.join(ROOT, "include", "dcsb200.h")).read()
    declared = set(re.findall(r"\b(dcsb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), name
    assert set(_capi.SYMBOLS) == declared


def test_sass_is_sm100(built):
    import shutil, subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", os.path.join(ROOT, "dcsexplorer_b200", "libdcsb200.so")],
                         capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_gain_helpers_match_oracle(built):
    from dcsexplorer_b200 import _capi
    from oracle import orc
    L, O = _capi.lib(), orc.lib()
    for v in range(-2, 300, 7):
        assert L.dcsb_master_multiplier(v) == O.dcso_master_multiplier(v)
    for s in range(-9000, 9000, 37):
        for osv in (0x9301, 0x9400):
            assert L.dcsb_level_multiplier(s, osv, 0xFF, 0) == O.dcso_level_multiplier(s, osv, 0xFF, 0)
            assert L.dcsb_level_multiplier(s, osv, 0x40, 1) == O.dcso_level_multiplier(s, osv, 0x40, 1)


def test_no_cpu_fallback(built):
    """Without a CUDA device the product must fail loudly, not decode on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import dcsexplorer_b200 as dx
    with pytest.raises(dx.DcsbError):
        dx.Context(0)


def test_null_and_bad_arguments_are_rejected_not_dereferenced(built):
    """Error behaviour of the boundary: integer status codes, no exceptions, no crashes on NULL
    handles or malformed arguments (host-only entry points and argument checks: no GPU needed)."""
    import ctypes as C
    import numpy as np
    import dcsexplorer_b200 as dx
    from dcsexplorer_b200 import _capi
    L = _capi.lib()
    E_ARG = dx._capi.E_ARG if hasattr(dx._capi, "E_ARG") else None
    out = C.c_void_p()
    assert L.dcsb_create(0, None) != dx.OK
    assert L.dcsb_create(-1, C.byref(out)) != dx.OK and not out.value
    L.dcsb_destroy(None)
    assert L.dcsb_last_error(None) == b"no context"
    assert L.dcsb_decode_streams(None, None, 0, None, None, None) != dx.OK
    assert L.dcsb_set_overlap(None, 1) != dx.OK and L.dcsb_set_pipeline(None, 0, 0) != dx.OK
    assert L.dcsb_render_timelines(None, None, None, 1, None, None, None) != dx.OK
    # ROM objects
    L.dcsb_rom_destroy(None)
    assert L.dcsb_rom_add(None, 2, None, 0) != dx.OK
    assert L.dcsb_rom_check(None) != 1
    assert L.dcsb_rom_track_info(None, 0, None) == 0
    assert L.dcsb_rom_decompile_track(None, 0, None, 0) == 0
    assert L.dcsb_rom_list_streams(None, None, 0) == 0
    assert not L.dcsb_rom_pointer(None, 0, None)
    rom = dx.Rom()
    img = np.zeros(1 << 19, dtype=np.uint8)
    for chip, n in ((1, 1 << 19), (10, 1 << 19), (2, 0), (2, 3000)):           # chip numbers 2..9, power-of-two sizes only
        assert L.dcsb_rom_add(rom._h, chip, img.ctypes.data, n) != dx.OK
    assert rom.check() != 1 and rom.info()["os"] == 0                          # nothing loaded: POST code 2 = U2 failed
    assert rom.track_info(0) is None and rom.decompile_track(0) == [] and rom.list_streams() == []
    assert rom.stream_bytes(0x123456, 4) == b""
    with pytest.raises(dx.DcsbError):
        dx.Rom(zip_path="/nonexistent/none.zip")
    rom.close()
    # players / batches without a context
    assert L.dcsb_player_create(None, None, C.byref(out)) != dx.OK
    L.dcsb_player_destroy(None)
    assert L.dcsb_player_render(None, 1, None) != dx.OK
    assert L.dcsb_player_host_bytes(None, None, 0) == 0
    assert L.dcsb_player_is_stream_playing(None, 0) == 0
    L.dcsb_batch_destroy(None)
    # the forward path: no context, no work; its size bound is pure host code (18 bytes of preamble + 526 per frame at most)
    assert L.dcsb_encode_streams(None, None, None, 1, None, None, 0, None, None) != dx.OK
    assert int(L.dcsb_encode_bound(0)) >= 18 and int(L.dcsb_encode_bound(240)) >= 18 + 526 and int(L.dcsb_encode_bound(241)) >= 18 + 2 * 526
    # stream partitioning is pure host code
    part, load = dx.partition_streams([10, 0, 7, 7, 3], 2)
    assert abs(int(load[0]) - int(load[1])) <= 3 and len(part) == 5 and set(int(x) for x in part) == {0, 1}


def test_public_header_is_plain_c(built, tmp_path):
    """include/dcsb200.h is the drop-in boundary: it must compile as C99 (no C++ / torch types)."""
    import shutil, subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("gcc not available")
    src = tmp_path / "hdr.c"
    src.write_text('#include "dcsb200.h"\nint main(void) { dcsb_opcode o; dcsb_stream_desc d; dcsb_result r; (void)o; (void)d; (void)r; return 0; }\n')
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_stream_extent_matches_oracle(built, golden):
    """dcsb_stream_extent (host side: the size of a stream handed over without one) against the oracle's
    frame walk on every golden stream and on fuzzer-made ones; damaged streams give 0 or a size within the data."""
    import ctypes as C
    import numpy as np
    import dcsfuzz
    from oracle import orc
    from dcsexplorer_b200 import _capi
    L = _capi.lib()
    items = [(it["stream"], it["os"]) for it in golden.items if not it["stop"]]
    items += [(d, os_) for os_, d, _ in dcsfuzz.corpus(seed=77, n_each=3, nframes=20)]
    checked = 0
    for d, os_ in items:
        nf = (d[0] << 8) | d[1]
        rc, bp, bt, stop = orc.scan(d, os_)
        buf = np.frombuffer(bytes(d) + bytes(32), dtype=np.uint8)
        ext = L.dcsb_stream_extent(buf.ctypes.data, os_)
        if rc == nf and nf and stop < 0:
            hdr_len = 1 if (os_ == 0x9301 and d[2] & 0x80) else 16
            assert ext == 2 + hdr_len + (int(bp[nf]) + 7) // 8, (hex(os_), nf)
            assert ext <= len(d)
            checked += 1
    assert checked >= 20
    assert L.dcsb_stream_extent(None, 0x9400) == 0
    z = np.zeros(64, dtype=np.uint8)
    assert L.dcsb_stream_extent(z.ctypes.data, 0x9400) == 0          # no frames
