"""Host replay of the device's Philox4x32-10 stream (mirrors philox4x32_10 in svinet_b200/csrc/svi_fa2_kernels.cuh).
TEST INFRASTRUCTURE: used to check the minibatches svi_fa2_draw / svi_fa2_run produce."""


def philox4x32_10(counter, key):
    c = [int(x) & 0xffffffff for x in counter]
    k0, k1 = int(key[0]) & 0xffffffff, int(key[1]) & 0xffffffff
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xffffffff, p1 & 0xffffffff, ((p0 >> 32) ^ c[3] ^ k1) & 0xffffffff,
             p0 & 0xffffffff]
        k0, k1 = (k0 + 0x9E3779B9) & 0xffffffff, (k1 + 0xBB67AE85) & 0xffffffff
    return c
