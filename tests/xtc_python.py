"""A pure-Python reader of XTC frames for the tests: the published libxdrf / xdrfile algorithm
(xdr3dfcoord) written the way the original is -- a three-integer bit-buffer state, byte arrays for
the mixed-radix numbers -- independently of ``cvpack_b200/csrc/xtc.cu`` (128-bit integers, one
bit cursor).  Test infrastructure only; slow (small files)."""

import struct

import numpy as np

MAGICINTS = [
    0, 0, 0, 0, 0, 0, 0, 0, 0, 8, 10, 12, 16, 20, 25, 32, 40, 50, 64,
    80, 101, 128, 161, 203, 256, 322, 406, 512, 645, 812, 1024, 1290,
    1625, 2048, 2580, 3250, 4096, 5060, 6501, 8192, 10321, 13003,
    16384, 20642, 26007, 32768, 41285, 52015, 65536, 82570, 104031,
    131072, 165140, 208063, 262144, 330280, 416127, 524287, 660561,
    832255, 1048576, 1321122, 1664510, 2097152, 2642245, 3329021,
    4194304, 5284491, 6658042, 8388607, 10568983, 13316085, 16777216,
]
FIRSTIDX = 9


def sizeofint(size):
    num, bits = 1, 0
    while size >= num and bits < 32:
        bits += 1
        num <<= 1
    return bits


def sizeofints(sizes):
    """Bits of the product of ``sizes``, multiplied out byte by byte like the original."""
    data = [1]
    for size in sizes:
        carry = 0
        for k, byte in enumerate(data):
            carry += byte * size
            data[k] = carry & 0xFF
            carry >>= 8
        while carry:
            data.append(carry & 0xFF)
            carry >>= 8
    num, bits = 1, 0
    while data[-1] >= num:
        bits += 1
        num *= 2
    return bits + 8 * (len(data) - 1)


class BitBuffer:
    def __init__(self, data):
        self.data = data
        self.cnt = 0
        self.lastbits = 0
        self.lastbyte = 0

    def receivebits(self, num_of_bits):
        mask = (1 << num_of_bits) - 1
        num = 0
        while num_of_bits >= 8:
            self.lastbyte = ((self.lastbyte << 8) | self.data[self.cnt]) & 0xFFFFFFFF
            self.cnt += 1
            num |= (self.lastbyte >> self.lastbits) << (num_of_bits - 8)
            num_of_bits -= 8
        if num_of_bits > 0:
            if self.lastbits < num_of_bits:
                self.lastbits += 8
                self.lastbyte = ((self.lastbyte << 8) | self.data[self.cnt]) & 0xFFFFFFFF
                self.cnt += 1
            self.lastbits -= num_of_bits
            num |= (self.lastbyte >> self.lastbits) & ((1 << num_of_bits) - 1)
        return num & mask

    def receiveints(self, num_of_bits, sizes):
        data = []
        while num_of_bits > 8:
            data.append(self.receivebits(8))
            num_of_bits -= 8
        if num_of_bits > 0:
            data.append(self.receivebits(num_of_bits))
        nums = [0, 0, 0]
        for i in (2, 1):
            num = 0
            for j in range(len(data) - 1, -1, -1):
                num = (num << 8) | data[j]
                quotient = num // sizes[i]
                data[j] = quotient
                num -= quotient * sizes[i]
            nums[i] = num
        data += [0, 0, 0, 0]
        nums[0] = data[0] | (data[1] << 8) | (data[2] << 16) | (data[3] << 24)
        return nums


def read_frames(path):
    """[(step, time, box[3][3], positions[N][3] float32), ...]"""
    blob = open(path, "rb").read()
    pos, frames = 0, []
    while pos < len(blob):
        magic, natoms, step = struct.unpack_from(">iii", blob, pos)
        assert magic == 1995
        (time,) = struct.unpack_from(">f", blob, pos + 12)
        box = np.array(struct.unpack_from(">9f", blob, pos + 16), dtype=np.float32).reshape(3, 3)
        pos += 52
        (lsize,) = struct.unpack_from(">i", blob, pos)
        assert lsize == natoms
        pos += 4
        if natoms <= 9:
            xyz = np.array(struct.unpack_from(f">{3 * natoms}f", blob, pos), dtype=np.float32).reshape(natoms, 3)
            pos += 12 * natoms
            frames.append((step, time, box, xyz))
            continue
        (precision,) = struct.unpack_from(">f", blob, pos)
        minint = struct.unpack_from(">3i", blob, pos + 4)
        maxint = struct.unpack_from(">3i", blob, pos + 16)
        smallidx, nbytes = struct.unpack_from(">ii", blob, pos + 28)
        pos += 36
        buf = BitBuffer(blob[pos:pos + nbytes] + b"\0\0\0\0")
        pos += (nbytes + 3) // 4 * 4
        sizeint = [maxint[k] - minint[k] + 1 for k in range(3)]
        if (sizeint[0] | sizeint[1] | sizeint[2]) > 0xFFFFFF:
            bitsizeint = [sizeofint(s) for s in sizeint]
            bitsize = 0
        else:
            bitsize = sizeofints(sizeint)
        smaller = MAGICINTS[max(FIRSTIDX, smallidx - 1)] // 2
        smallnum = MAGICINTS[smallidx] // 2
        sizesmall = [MAGICINTS[smallidx]] * 3
        inv_precision = np.float32(1.0) / np.float32(precision)
        out, run, i = [], 0, 0
        while i < natoms:
            if bitsize == 0:
                this = [buf.receivebits(bitsizeint[k]) for k in range(3)]
            else:
                this = buf.receiveints(bitsize, sizeint)
            i += 1
            this = [this[k] + minint[k] for k in range(3)]
            prev = list(this)
            is_smaller = 0
            if buf.receivebits(1) == 1:
                run = buf.receivebits(5)
                is_smaller = run % 3
                run -= is_smaller
                is_smaller -= 1
            if run > 0:
                for k in range(0, run, 3):
                    this = buf.receiveints(smallidx, sizesmall)
                    i += 1
                    this = [this[c] + prev[c] - smallnum for c in range(3)]
                    if k == 0:
                        this, prev = prev, this
                        out.append(prev)
                    else:
                        prev = list(this)
                    out.append(this)
            else:
                out.append(this)
            smallidx += is_smaller
            if is_smaller < 0:
                smallnum = smaller
                smaller = MAGICINTS[smallidx - 1] // 2 if smallidx > FIRSTIDX else 0
            elif is_smaller > 0:
                smaller = smallnum
                smallnum = MAGICINTS[smallidx] // 2
            sizesmall = [MAGICINTS[smallidx]] * 3
        xyz = np.array(out, dtype=np.int64).astype(np.float32) * inv_precision
        frames.append((step, time, box, xyz.astype(np.float32)))
    return frames
