"""Pin oracle/philox.py to the published Philox4x32-10 through the Random123 known-answer vectors."""
import numpy as np

from oracle.philox import normal2, philox4x32_10


def test_random123_known_answers():
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == (
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)


def test_normals_are_standard_and_keyed():
    z = np.array([normal2(7, 3, e) for e in range(20000)])
    assert abs(z.mean()) < 0.02 and abs(z.std() - 1) < 0.02 and abs(np.mean(z[:, 0] * z[:, 1])) < 0.02
    assert normal2(7, 3, 5) == normal2(7, 3, 5) and normal2(7, 3, 5) != normal2(7, 4, 5) and normal2(7, 3, 5) != normal2(8, 3, 5)
