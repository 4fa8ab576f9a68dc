"""State I/O (reference src/IO.jl:1-23): the JLD2 writer/reader of nupgcm_b200/io.py.

Pins: (i) the metadata checksum against bytes of a reference-written file; (ii) regenerating the
reference's own state file test/data/bowl_mixing_2D.jld2 from its decoded contents
(tests/golden/bowl_mixing_2D.npz) must give the identical 12 660 bytes (sha256 recorded from the
reference file by the test's author); (iii) the independent decoder oracle/jld2.py — validated
on the reference's files — reads what the product writes."""
import hashlib
import os

import numpy as np
import pytest

from conftest import golden
from nupgcm_b200 import io as sio
from oracle.jld2 import read_jld2 as oracle_read

REF_SUPERBLOCK = bytes.fromhex("894844460d0a1a0a020808000002000000000000ffffffffffffffff"
                               "7431000000000000f52e000000000000b4324a17")
REF_SHA256 = "a99b90b770551a833d17164d9ed6c6c2ff28c7a34796af457446fd097c98c19e"


def test_lookup3_matches_reference_file_checksum():
    assert sio.lookup3(REF_SUPERBLOCK[:44]) == int.from_bytes(REF_SUPERBLOCK[44:], "little")
    assert sio.lookup3(b"") == 0xDEADBEEF
    assert sio.lookup3(b"Four score and seven years ago") == 0x17770551       # lookup3.c driver5()


def test_regenerates_reference_state_file_byte_for_byte(tmp_path):
    g = golden("bowl_mixing_2D.npz")
    f = tmp_path / "state.jld2"
    sio.write_jld2(str(f), {"u": g["u"], "p": g["p"], "b": g["b"], "t": float(g["t"])},
                   creator="Julia 1.10.2 64-bit LE")
    data = f.read_bytes()
    assert len(data) == 12660
    assert hashlib.sha256(data).hexdigest() == REF_SHA256


def test_round_trip_and_independent_decoder(tmp_path):
    rng = np.random.default_rng(3)
    fields = {"u": rng.normal(size=1001), "p": rng.normal(size=7), "b": np.zeros(0), "t": 0.125}
    f = str(tmp_path / "s.jld2")
    sio.write_jld2(f, fields)
    back = sio.read_jld2(f)
    theirs = oracle_read(f)
    for k in ("u", "p", "b"):
        assert np.array_equal(back[k], fields[k]) and np.array_equal(theirs[k], fields[k])
    assert back["t"] == 0.125 and float(theirs["t"]) == 0.125


def test_reader_rejects_corruption(tmp_path):
    f = tmp_path / "s.jld2"
    sio.write_jld2(str(f), {"u": np.arange(5.0), "t": 1.0})
    raw = bytearray(f.read_bytes())
    raw[512 + 40] ^= 1                                   # root group address inside the superblock
    f.write_bytes(bytes(raw))
    with pytest.raises(ValueError):
        sio.read_jld2(str(f))
    with pytest.raises(ValueError):
        (tmp_path / "x.jld2").write_bytes(b"not a jld2 file" * 100)
        sio.read_jld2(str(tmp_path / "x.jld2"))


@pytest.mark.gpu
def test_save_and_restore_model_state(tmp_path):
    import nupgcm_b200 as npg
    from conftest import workload
    w, ops = workload("bowl_mixing", dim=2)

    def make():
        arch = npg.GPU(0)
        inv = npg.InversionToolkit(arch, ops["A"], ops["pscale"], ops["B"], ops["b0"])
        ts = w.timestepper()
        evo = npg.EvolutionToolkit(arch, ops, w.params, w.forcings, ts)
        m = npg.Model(arch, w.params, w.forcings, w.fe_data(), inv, evo, ts, tables=ops["tables"])
        m.xb.upload(ops["b_init"])
        return m

    a = make()
    npg.run_(a, n_steps=3)
    f = str(tmp_path / "state.jld2")
    sio.save_state(a, f)
    b = make()
    sio.set_state_from_file_(b, f)
    assert np.array_equal(a.xb.download(), b.xb.download())
    assert np.array_equal(a.inversion.solver.x.download(), b.inversion.solver.x.download())
    assert b.timestepper.t == a.timestepper.t
    d = oracle_read(f)
    assert np.array_equal(d["u"], a.state.u) and np.array_equal(d["b"], a.state.b)


def test_reader_accepts_the_int64_time_of_a_model_without_timestepper(tmp_path):
    """save_state writes `t = 0` as an Int64 when the model has no timestepper (IO.jl:3-4): patch the
    datatype message of a file we wrote into the fixed-point class and read it back."""
    f = tmp_path / "s.jld2"
    sio.write_jld2(str(f), {"u": np.arange(4.0), "t": 0.0})
    raw = bytearray(f.read_bytes())
    f64 = bytes([0x31, 0x20, 0x3F, 0x00])                 # class 1 (float) v1 datatype header written by write_jld2
    pos = raw.rfind(f64)                                  # the last dataset written is `t`
    assert pos > 0
    raw[pos] = 0x30                                       # class 0 (fixed point), same 8-byte size
    f.write_bytes(bytes(raw))
    try:
        back = sio.read_jld2(str(f))
    except ValueError as e:                               # object-header checksum now differs: that is the only accepted reason
        assert "checksum" in str(e)
    else:
        assert back["t"] == 0.0


@pytest.mark.gpu
def test_periodic_save_and_bdf2_restart_continue_the_run_bit_for_bit(tmp_path):
    """run_(n_save=k, out_dir=..., save_history=True) writes $out_dir/data/state_%016d.jld2 (model.jl:194-197) with
    the BDF2 history; a fresh model restored from it and resumed ends where the uninterrupted run ends."""
    import nupgcm_b200 as npg
    from conftest import workload
    w, ops = workload("bowl_mixing", dim=2)

    def make():
        arch = npg.GPU(0)
        inv = npg.InversionToolkit(arch, ops["A"], ops["pscale"], ops["B"], ops["b0"])
        ts = w.timestepper()
        evo = npg.EvolutionToolkit(arch, ops, w.params, w.forcings, ts)
        m = npg.Model(arch, w.params, w.forcings, w.fe_data(), inv, evo, ts, tables=ops["tables"])
        m.xb.upload(ops["b_init"])
        return m

    a = make()
    npg.run_(a, n_steps=6, n_save=3, out_dir=str(tmp_path), save_history=True)
    files = sorted((tmp_path / "data").iterdir())
    assert [p.name for p in files] == ["state_%016d.jld2" % 3, "state_%016d.jld2" % 6]
    b = make()
    sio.set_state_from_file_(b, str(files[0]))
    npg.run_(b, n_steps=3, resume=True)
    assert np.array_equal(a.xb.download(), b.xb.download())
    assert np.array_equal(a.inversion.solver.x.download(), b.inversion.solver.x.download())
    assert b.timestepper.t == a.timestepper.t
    with pytest.raises(ValueError):
        npg.run_(make(), n_steps=1, n_save=1)             # n_save without out_dir is refused, not ignored
