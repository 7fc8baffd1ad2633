"""host/pdeflate.c: the multi-threaded container stage must emit ONE valid zlib
stream that inflates to exactly its input, whatever the thread count, block size and write pattern.
The reference's own reader is `inflate` (C/decoder.c:213-227); Python's zlib is the same library."""
import os
import subprocess
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("pdeflate") / "pdeflate_driver")
    subprocess.check_call(["gcc", "-O2", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "pdeflate_driver.c"),
                           os.path.join(ROOT, "host", "pdeflate.c"), "-lz"])
    return exe


def run(exe, data, level=9, threads=4, block=65536, chunk=100000):
    p = subprocess.run([exe, str(level), str(threads), str(block), str(chunk)], input=data, capture_output=True, check=True)
    n_in, n_out = (int(x) for x in p.stderr.split())
    assert n_in == len(data) and n_out == len(p.stdout)
    return p.stdout


def eg_like(n, seed):
    """Bytes shaped like an Exp-Golomb stream of a natural clip: long runs of 0xff with sparse other bytes."""
    rng = np.random.default_rng(seed)
    a = np.full(n, 0xFF, np.uint8)
    idx = rng.random(n) < 0.15
    a[idx] = rng.integers(0, 256, int(idx.sum()), dtype=np.uint8)
    return a.tobytes()


@pytest.mark.parametrize("n", [0, 1, 65535, 65536, 65537, 3 * 65536, 1_000_003])
def test_round_trip_sizes(driver, n):
    data = eg_like(n, n)
    z = run(driver, data)
    assert z[0] == 0x78 and (z[0] * 256 + z[1]) % 31 == 0          # RFC 1950 header
    d = zlib.decompressobj()
    assert d.decompress(z) == data and d.eof and d.unused_data == b""   # one complete stream, Adler-32 verified


@pytest.mark.parametrize("threads,block,chunk", [(1, 32768, 7), (2, 1024, 4096), (16, 262144, 1 << 20), (5, 100000, 99999)])
def test_threads_blocks_and_write_patterns(driver, threads, block, chunk):
    data = eg_like(700_001, 3) + os.urandom(50_000) + bytes(200_000)
    z = run(driver, data, 9, threads, block, chunk)
    assert zlib.decompress(z) == data


def test_levels_and_ratio_close_to_single_stream(driver):
    data = eg_like(2_000_000, 5)
    for level in (1, 6, 9):
        z = run(driver, data, level, 8, 262144, 1 << 20)
        assert zlib.decompress(z) == data
        ref = zlib.compress(data, level)
        assert len(z) < 1.03 * len(ref) + 64      # the cuts cost little thanks to the carried dictionary


def test_reference_cli_reads_our_container(driver, tmp_path):
    """The reference's own decoder (unmodified C/decoder.c + inflate, built into oracle/_ref/codec_ref) must read a file
    whose container was written by pdeflate: Exp-Golomb stream from the oracle, our container, the reference's decode."""
    import importlib
    import sys
    ref = os.path.join(ROOT, "oracle", "_ref", "codec_ref")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/codec_ref not built (needs /root/reference at build time)")
    sys.path.insert(0, ROOT)
    O = importlib.import_module("oracle.oracle")
    O.build()
    synth = importlib.import_module("3ddctvideoencoding_b200.synth")
    W, H, F = 64, 48, 24
    clip = synth.natural(W, H, F, 7)
    stream, nbits = O.encode_u8(clip, 8, 1)             # C rounding flavour, as codec_ref itself would produce
    z = run(driver, stream.tobytes(), 9, 4, 1024, 777)  # many small blocks: every cut is exercised
    (tmp_path / "a.dct").write_bytes(z)
    p = subprocess.run([ref, "decode", str(tmp_path / "a.dct"), str(tmp_path / "a.out"), str(W), str(H), str(F), "1"],
                       capture_output=True, cwd=os.path.dirname(ref), timeout=300)
    assert p.returncode == 0, p.stdout[-500:]
    got = np.fromfile(str(tmp_path / "a.out"), np.uint8).reshape(F, H, W)
    want = O.decode_u8(stream, W, H, F, 8)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 1       # float kernels of the reference vs the fp64 oracle
