"""ZSTD_compress_usingDict path of the oracle (BASELINE config 5 shape: small records + shared dictionary).
Frames must decode with the reference's ZSTD_decompress_usingDict; raw-content dictionaries must stay
within +-0.5 % of the reference's size; zstd-format dictionaries are used for their content only (the
entropy-table reuse the reference gets from them is a documented gap)."""
import pytest

import zref

needs_ref = pytest.mark.skipif(not zref.have_ref(), reason="oracle/_ref/libzstd_ref.so not built")
REC = 1024


def records(n, seed, p=0.5):
    data = zref.synthetic(REC * n, seed, p)
    return [data[i * REC:(i + 1) * REC] for i in range(n)]


@needs_ref
@pytest.mark.parametrize("dict_name", ["zdict-16k-synthetic-seed77", "http-dict-missing-symbols", "zero-weight-dict"])
def test_zstd_format_dictionaries_roundtrip(dict_name):
    d = zref.golden_input(dict_name)
    tot_o = tot_r = 0
    srcs = records(200, 5) + [b"", b"a", zref.golden_input("http"), zref.synthetic(300_000, 8)]
    for src in srcs:
        f = zref.oracle_compress_using_dict(src, d, 1)
        assert zref.ref_decompress_using_dict(f, d, len(src)) == src
        assert f[4] & 3, "dictID must be present in the frame header for a zstd-format dictionary"
        tot_o += len(f)
        tot_r += len(zref.ref_compress_using_dict(src, d, 1))
    assert tot_o < tot_r * 1.06          # content-only use of the dictionary: a few % behind the reference


@needs_ref
def test_raw_content_dictionary_size_parity():
    d = zref.synthetic(32 << 10, 123, 0.5)                 # no magic number -> raw content (zstd_compress.c:5143-5148)
    tot_o = tot_r = 0
    for src in records(400, 6):
        f = zref.oracle_compress_using_dict(src, d, 1)
        assert zref.ref_decompress_using_dict(f, d, len(src)) == src
        assert (f[4] & 3) == 0                              # no dictID for raw content (lib/zstd.h:185-186)
        tot_o += len(f)
        tot_r += len(zref.ref_compress_using_dict(src, d, 1))
    assert abs(tot_o - tot_r) / tot_r <= 0.005


@needs_ref
def test_dictionary_content_is_actually_used():
    """A record that is a verbatim slice of the dictionary must compress to almost nothing."""
    d = zref.synthetic(32 << 10, 321, 0.1)
    src = d[5000:6024]
    with_dict = zref.oracle_compress_using_dict(src, d, 1)
    without = zref.oracle_compress(src, 1)
    assert zref.ref_decompress_using_dict(with_dict, d, len(src)) == src
    assert len(with_dict) < 100 < len(without)


@needs_ref
def test_short_and_corrupted_dictionaries():
    src = zref.synthetic(5000, 1)
    assert zref.oracle_compress_using_dict(src, b"1234567", 1) == zref.oracle_compress(src, 1)      # < 8 bytes: ignored
    bad = bytes.fromhex("37a430ec01000000") + bytes(40)
    with pytest.raises(RuntimeError, match="30"):                                                     # dictionary_corrupted
        zref.oracle_compress_using_dict(src, bad, 1)
