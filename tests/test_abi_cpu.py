"""CPU: checks of libtezgpu that need no device -- the GF(2) algebra of the tiled, parallel CRC32 (host emulation on
the same tables the kernels use) against zlib."""
import random
import zlib

from tez_b200 import _lib


def test_tiled_crc_algebra_matches_zlib():
    L = _lib.load()
    rng = random.Random(1)
    for n in [0, 1, 3, 4, 5, 15, 16, 17, 63, 64, 1000, 1027, 4096 * 3 + 5, 24 * 1024 - 7, 24 * 1024 + 9, 100000, 1 << 20]:
        d = bytes(rng.getrandbits(8) for _ in range(n))
        for piece in (64, 1040, 24 * 1024):
            for lead in (0, 1, 7, 12, 15):
                assert L.tezgpu_debug_crc_emulate(d, n, piece, lead) == zlib.crc32(d), (n, piece, lead)
