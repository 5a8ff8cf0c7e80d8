import numpy as np


def test_torch_generator_matches_oracle_generator(oracle):
    from nvjpeg_imagecompressor_b200.synth import synth
    for (W, H, seed, amp) in ((70, 50, 0, 8), (33, 100, 5, 4), (8320, 3, 0, 8)):
        a = synth(W, H, seed, amp, device="cpu").numpy()
        assert np.array_equal(a, oracle.synth(W, H, seed, amp))
