"""Output containers: the 44-byte WAV file DCSExplorer's extractor writes
(DCSExplorer.cpp:1686-1712) and the raw "DCSa" stream container (:1831-1871; reader
DCSEncoder.cpp:369-399).  Host side, no GPU."""
import struct
import wave
import numpy as np
import pytest
import dcsfuzz


def test_wav_writer_layout_and_roundtrip(built, tmp_path):
    import dcsexplorer_b200 as dx
    pcm = (np.arange(-1200, 1200, dtype=np.int32) * 27 % 65536 - 32768).astype(np.int16)
    p = tmp_path / "x.wav"
    dx.write_wav(p, pcm)
    raw = p.read_bytes()
    assert len(raw) == 44 + pcm.size * 2
    # field by field, as the reference's extractor fills them
    assert raw[0:4] == b"RIFF" and raw[8:16] == b"WAVEfmt " and raw[36:40] == b"data"
    riff, fmtlen, fmt, ch, rate, bps, align, bits = struct.unpack("<I", raw[4:8]) + struct.unpack("<IHHIIHH", raw[16:36])
    assert (riff, fmtlen, fmt, ch, rate, bps, align, bits) == (pcm.size * 2 + 36, 16, 1, 1, 31250, 62500, 2, 16)
    assert struct.unpack("<I", raw[40:44])[0] == pcm.size * 2
    with wave.open(str(p)) as w:                 # and a standard reader accepts it
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 31250, pcm.size)
        assert np.array_equal(np.frombuffer(w.readframes(pcm.size), dtype="<i2"), pcm)
    dx.write_wav(tmp_path / "empty.wav", np.zeros(0, dtype=np.int16))
    assert (tmp_path / "empty.wav").stat().st_size == 44
    with pytest.raises(dx.DcsbError):
        dx.write_wav(tmp_path / "no_such_dir" / "x.wav", pcm)


@pytest.mark.parametrize("osv,tag", [(0x9400, b"\x94\x00"), (0x9500, b"\x94\x00"), (0x9302, b"\x93\x02"), (0x9301, b"\x93\x01")])
def test_dcs_container_roundtrip(built, tmp_path, osv, tag):
    import dcsexplorer_b200 as dx
    rng = np.random.default_rng(osv)
    d = dcsfuzz.fuzz94(rng, 9) if osv >= 0x9400 else dcsfuzz.fuzz93(rng, 9, type1=0)
    p = tmp_path / "s.dcs"
    dx.write_dcs_file(p, osv, d)
    raw = p.read_bytes()
    assert raw[:4] == b"DCSa" and raw[4:6] == tag and raw[6:10] == b"\x00\x01\x7a\x12" and raw[10:32] == bytes(22)
    assert struct.unpack(">I", raw[32:36])[0] == len(d) and raw[36:] == d
    v, back = dx.read_dcs_file(p)
    assert back == d and v == ((tag[0] << 8) | tag[1])
    (tmp_path / "bad.dcs").write_bytes(b"RIFF" + raw[4:])
    with pytest.raises(dx.DcsbError):
        dx.read_dcs_file(tmp_path / "bad.dcs")
    with pytest.raises(dx.DcsbError):
        dx.read_dcs_file(tmp_path / "missing.dcs")


@pytest.mark.parametrize("osv,fmt", [(0x9400, 0x9400), (0x9500, 0x9400), (0x9302, 0x9302), (0x9301, 0x9301)])
def test_dcs_container_is_read_by_the_reference(built, tmp_path, osv, fmt):
    """A file written by dcsb_write_dcs_file goes through the reference's own reader (DCSEncoder::IsDCSFile /
    EncodeDCSFile, DCSEncoder.cpp:358-519, compiled unmodified into oracle/_ref): recognised, the format version
    it names is the stream's, and the stream bytes and frame count it takes from the file are the ones written."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    import dcsexplorer_b200 as dx
    rng = np.random.default_rng(osv + 1)
    d = dcsfuzz.fuzz94(rng, 23, type1=1) if osv >= 0x9400 else dcsfuzz.fuzz93(rng, 23, type1=0)
    p = tmp_path / "r.dcs"
    dx.write_dcs_file(p, osv, d)
    got = ref.read_dcs_file(p)
    assert got is not None, "the reference does not recognise the file"
    assert got[0] == fmt and got[1] == d and got[2] == ((d[0] << 8) | d[1])
    # and a WAV file is not mistaken for one
    dx.write_wav(tmp_path / "x.wav", np.zeros(100, dtype=np.int16))
    assert ref.read_dcs_file(tmp_path / "x.wav") is None
