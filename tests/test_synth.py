"""CPU: the torch generators used by bench.py / GPU tests are bit-identical to the oracle's C generators."""
import numpy as np
import torch

from oracle import tez_oracle as O
from tez_b200 import synth


def test_gen_c2_matches_oracle():
    for first, n, seed in [(0, 1000, 2), (12345, 777, 4), (10 ** 9, 64, 11)]:
        a = synth.gen_c2(first, n, seed=seed).numpy()
        b = O.gen_c2(first, n, seed=seed)
        assert np.array_equal(a, b)


def test_hash_partition_matches_oracle():
    kv = O.gen_c2(0, 500, seed=2).reshape(500, 80)
    keys = torch.from_numpy(kv[:, :16].copy())
    got = synth.hash_partition(keys, 64).numpy()
    exp = np.array([O.partition_of(O.CMP_BYTES, kv[i, :16].tobytes(), 64) for i in range(500)])
    assert np.array_equal(got, exp)


def test_threaded_c_generator_gives_the_same_bytes():
    """bench.py's CPU arm fills its sample with the oracle's C generator on many threads (slices are independent)."""
    import numpy as np
    from oracle import tez_oracle as O
    n = (1 << 20) + 12345
    assert np.array_equal(O.gen_c2(7, n, seed=2, threads=5), O.gen_c2(7, n, seed=2))
