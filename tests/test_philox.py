"""Known-answer vectors of Philox4x32-10 (Salmon et al., Random123 kat_vectors) for the host replay used to check
the device's minibatch stream (tests/philox_ref.py == philox4x32_10 in svinet_b200/csrc/svi_fa2_kernels.cuh;
test_gpu_fa2.py::test_device_draw_and_run compares the device's draws with this replay)."""
from philox_ref import philox4x32_10

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        assert tuple(philox4x32_10(ctr, key)) == want
