// Trajectory ingestion, native side: GROMACS XTC frames ->
// [count][N][3] float nm in a caller-owned (pinned) buffer, and the matching writer.
//
// XTC is the compressed format of libxdrf / xdrfile (GROMACS; F. van Hoesel's xdr3dfcoord): every
// frame is an XDR (big-endian) header -- magic 1995, atom count, step, time, 3x3 box in nm -- and
// a coordinate block in which positions are rounded to integers (x * precision), packed either as
// one mixed-radix number per atom (radices = extent of the frame's bounding box) or, for runs of
// atoms close to their predecessor (waters), as small mixed-radix numbers relative to it with an
// adaptive radix taken from a table of ~2^(k/3) ("magic ints").  Neither GROMACS nor xdrfile is in
// this image and the reference ships no XTC file, so the bit layout below restates the published
// algorithm; tests/test_trajectory.py pins the header and the uncompressed (<= 9 atoms) form to
// hand-assembled bytes and the compressed form to an independent pure-Python decoder.
//
// Frames have variable size: cvp_xtc_open walks the headers once and keeps an offset index, so
// chunks are random-access and host threads decode different frames at the same time, each
// straight into its slot of the destination.  Host code only -- no kernel is launched here.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

struct cvp_xtc {
  int fd = -1;
  const unsigned char* map = nullptr;
  size_t size = 0;
  int64_t num_atoms = 0;
  std::vector<int64_t> offsets;  // start of every frame
  std::vector<int32_t> steps;
  std::vector<float> times;
  float precision = 0.0f;
};

namespace cvp {
namespace {

constexpr int XTC_MAGIC = 1995;
constexpr int FIRSTIDX = 9;
// magicints[k] ~ 2^(k/3): three integers below magicints[k] fit k bits
const int magicints[] = {0,       0,       0,       0,       0,       0,       0,       0,        0,        8,
                         10,      12,      16,      20,      25,      32,      40,      50,       64,       80,
                         101,     128,     161,     203,     256,     322,     406,     512,      645,      812,
                         1024,    1290,    1625,    2048,    2580,    3250,    4096,    5060,     6501,     8192,
                         10321,   13003,   16384,   20642,   26007,   32768,   41285,   52015,    65536,    82570,
                         104031,  131072,  165140,  208063,  262144,  330280,  416127,  524287,   660561,   832255,
                         1048576, 1321122, 1664510, 2097152, 2642245, 3329021, 4194304, 5284491,  6658042,  8388607,
                         10568983, 13316085, 16777216};
constexpr int LASTIDX = (int)(sizeof(magicints) / sizeof(*magicints));

typedef unsigned __int128 u128;

uint32_t be32(const unsigned char* p) {
  return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}
float be_float(const unsigned char* p) {
  const uint32_t v = be32(p);
  float f;
  std::memcpy(&f, &v, 4);
  return f;
}
void put_be32(std::vector<unsigned char>& out, uint32_t v) {
  out.push_back((unsigned char)(v >> 24));
  out.push_back((unsigned char)(v >> 16));
  out.push_back((unsigned char)(v >> 8));
  out.push_back((unsigned char)v);
}
void put_be_float(std::vector<unsigned char>& out, float f) {
  uint32_t v;
  std::memcpy(&v, &f, 4);
  put_be32(out, v);
}

// bits needed for values 0 .. size
int bits_for(unsigned size) {
  unsigned num = 1;
  int bits = 0;
  while (size >= num && bits < 32) {
    ++bits;
    num <<= 1;
  }
  return bits;
}

// bits needed for the mixed-radix number with radices sizes[0..2]
int bits_for_product(const unsigned sizes[3]) {
  u128 product = (u128)sizes[0] * sizes[1] * sizes[2];  // < 2^96
  // the packed value is at most product - 1; the format sizes it from the byte-wise product itself
  int bits = 0;
  while (product != 0) {
    ++bits;
    product >>= 1;
  }
  return bits;
}

// MSB-first bit stream over a byte range (reads past the end return zero bits and set `overrun`)
struct BitReader {
  const unsigned char* data;
  size_t bytes;
  size_t pos = 0;  // in bits
  bool overrun = false;
  unsigned get(int nbits) {  // nbits <= 32
    uint64_t value = 0;
    int left = nbits;
    while (left > 0) {
      const size_t byte = pos >> 3;
      const int offset = (int)(pos & 7);
      const int take = std::min(left, 8 - offset);
      unsigned cur = 0;
      if (byte < bytes)
        cur = data[byte];
      else
        overrun = true;
      const unsigned piece = (cur >> (8 - offset - take)) & ((1u << take) - 1u);
      value = (value << take) | piece;
      pos += take;
      left -= take;
    }
    return (unsigned)value;
  }
  // a number sent as whole bytes, least significant byte first, the last one `nbits % 8` wide
  u128 get_bytes_le(int nbits) {
    u128 value = 0;
    int shift = 0;
    while (nbits > 8) {
      value |= (u128)get(8) << shift;
      shift += 8;
      nbits -= 8;
    }
    if (nbits > 0) value |= (u128)get(nbits) << shift;
    return value;
  }
  void get_ints(int nbits, const unsigned sizes[3], int out[3]) {
    u128 v = get_bytes_le(nbits);
    out[2] = (int)(v % sizes[2]);
    v /= sizes[2];
    out[1] = (int)(v % sizes[1]);
    v /= sizes[1];
    out[0] = (int)(unsigned)v;
  }
};

struct BitWriter {
  std::vector<unsigned char> data;
  int free_bits = 0;  // unused low bits of the last byte
  void put(int nbits, unsigned value) {
    while (nbits > 0) {
      if (free_bits == 0) {
        data.push_back(0);
        free_bits = 8;
      }
      const int take = std::min(nbits, free_bits);
      const unsigned piece = (value >> (nbits - take)) & ((1u << take) - 1u);
      data.back() |= (unsigned char)(piece << (free_bits - take));
      free_bits -= take;
      nbits -= take;
    }
  }
  void put_bytes_le(int nbits, u128 value) {
    while (nbits > 8) {
      put(8, (unsigned)(value & 0xff));
      value >>= 8;
      nbits -= 8;
    }
    if (nbits > 0) put(nbits, (unsigned)(value & ((1u << nbits) - 1u)));
  }
  void put_ints(int nbits, const unsigned sizes[3], const int in[3]) {
    const u128 v = ((u128)(unsigned)in[0] * sizes[1] + (unsigned)in[1]) * sizes[2] + (unsigned)in[2];
    put_bytes_le(nbits, v);
  }
};

// size of the coordinate block that starts at p (after the 3x3 box), or -1
int64_t coordinate_block_bytes(const unsigned char* p, size_t available, int64_t natoms) {
  if (available < 4 || (int64_t)be32(p) != natoms) return -1;
  if (natoms <= 9) return 4 + 12 * natoms;
  if (available < 40) return -1;
  const int64_t packed = (int32_t)be32(p + 36);
  if (packed < 0) return -1;
  return 40 + ((packed + 3) & ~(int64_t)3);
}

// one frame's coordinate block -> out[N][3] (nm); returns false on a malformed block
bool decode_coordinates(const unsigned char* p, int64_t natoms, float* out, float* precision_out) {
  if (natoms <= 9) {
    for (int64_t i = 0; i < 3 * natoms; ++i) out[i] = be_float(p + 4 + 4 * i);
    if (precision_out) *precision_out = 0.0f;
    return true;
  }
  const float precision = be_float(p + 4);
  if (precision_out) *precision_out = precision;
  int minint[3], maxint[3];
  for (int k = 0; k < 3; ++k) {
    minint[k] = (int32_t)be32(p + 8 + 4 * k);
    maxint[k] = (int32_t)be32(p + 20 + 4 * k);
  }
  int smallidx = (int32_t)be32(p + 32);
  const int64_t packed = (int32_t)be32(p + 36);
  if (smallidx < FIRSTIDX || smallidx >= LASTIDX || !(precision > 0.0f)) return false;
  unsigned sizeint[3];
  int bitsizeint[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    if (maxint[k] < minint[k]) return false;
    sizeint[k] = (unsigned)((int64_t)maxint[k] - minint[k] + 1);
  }
  int bitsize;
  if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffff) {
    for (int k = 0; k < 3; ++k) bitsizeint[k] = bits_for(sizeint[k]);
    bitsize = 0;
  } else {
    bitsize = bits_for_product(sizeint);
  }
  int smaller = magicints[std::max(FIRSTIDX, smallidx - 1)] / 2;
  int smallnum = magicints[smallidx] / 2;
  unsigned sizesmall[3] = {(unsigned)magicints[smallidx], (unsigned)magicints[smallidx], (unsigned)magicints[smallidx]};
  const float inv_precision = 1.0f / precision;
  BitReader bits{p + 40, (size_t)packed};
  int run = 0;
  int64_t i = 0;
  float* dst = out;
  while (i < natoms) {
    int cur[3], prev[3];
    if (bitsize == 0) {
      for (int k = 0; k < 3; ++k) cur[k] = (int)bits.get(bitsizeint[k]);
    } else {
      bits.get_ints(bitsize, sizeint, cur);
    }
    ++i;
    for (int k = 0; k < 3; ++k) {
      cur[k] += minint[k];
      prev[k] = cur[k];
    }
    int is_smaller = 0;
    if (bits.get(1) == 1) {
      run = (int)bits.get(5);
      is_smaller = run % 3;
      run -= is_smaller;
      --is_smaller;
    }
    if (run > 0) {
      if (i + run / 3 > natoms) return false;
      for (int k = 0; k < run; k += 3) {
        bits.get_ints(smallidx, sizesmall, cur);
        ++i;
        for (int c = 0; c < 3; ++c) cur[c] += prev[c] - smallnum;
        if (k == 0) {
          // the writer swapped the first two atoms of the run (water: O after H compresses better)
          for (int c = 0; c < 3; ++c) std::swap(cur[c], prev[c]);
          for (int c = 0; c < 3; ++c) *dst++ = (float)prev[c] * inv_precision;
        } else {
          for (int c = 0; c < 3; ++c) prev[c] = cur[c];
        }
        for (int c = 0; c < 3; ++c) *dst++ = (float)cur[c] * inv_precision;
      }
    } else {
      for (int c = 0; c < 3; ++c) *dst++ = (float)cur[c] * inv_precision;
    }
    smallidx += is_smaller;
    if (smallidx < FIRSTIDX || smallidx >= LASTIDX) return false;
    if (is_smaller < 0) {
      smallnum = smaller;
      smaller = smallidx > FIRSTIDX ? magicints[smallidx - 1] / 2 : 0;
    } else if (is_smaller > 0) {
      smaller = smallnum;
      smallnum = magicints[smallidx] / 2;
    }
    sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
    if (bits.overrun) return false;
  }
  return true;
}

// the writer's side of the same block; false when a coordinate does not fit an int at this precision
bool encode_coordinates(const float* x, int64_t natoms, float precision, std::vector<unsigned char>& out) {
  put_be32(out, (uint32_t)natoms);
  if (natoms <= 9) {
    for (int64_t i = 0; i < 3 * natoms; ++i) put_be_float(out, x[i]);
    return true;
  }
  if (!(precision > 0.0f)) precision = 1000.0f;
  put_be_float(out, precision);
  std::vector<int> ints((size_t)natoms * 3 + 3, 0);
  int minint[3] = {INT_MAX, INT_MAX, INT_MAX}, maxint[3] = {INT_MIN, INT_MIN, INT_MIN};
  int mindiff = INT_MAX;
  int old[3] = {0, 0, 0};
  for (int64_t i = 0; i < natoms; ++i) {
    int v[3];
    for (int k = 0; k < 3; ++k) {
      const float f = x[3 * i + k];
      const float lf = f >= 0.0f ? f * precision + 0.5f : f * precision - 0.5f;
      if (!(std::fabs(lf) <= (float)(INT_MAX - 2))) return false;
      v[k] = (int)lf;
      minint[k] = std::min(minint[k], v[k]);
      maxint[k] = std::max(maxint[k], v[k]);
      ints[3 * i + k] = v[k];
    }
    const int64_t diff = std::llabs((int64_t)old[0] - v[0]) + std::llabs((int64_t)old[1] - v[1]) +
                         std::llabs((int64_t)old[2] - v[2]);
    if (diff < mindiff && i > 0) mindiff = (int)diff;
    for (int k = 0; k < 3; ++k) old[k] = v[k];
  }
  for (int k = 0; k < 3; ++k) put_be32(out, (uint32_t)minint[k]);
  for (int k = 0; k < 3; ++k) put_be32(out, (uint32_t)maxint[k]);
  unsigned sizeint[3];
  int bitsizeint[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    if ((int64_t)maxint[k] - minint[k] >= (int64_t)INT_MAX - 2) return false;
    sizeint[k] = (unsigned)(maxint[k] - minint[k] + 1);
  }
  int bitsize;
  if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffff) {
    for (int k = 0; k < 3; ++k) bitsizeint[k] = bits_for(sizeint[k]);
    bitsize = 0;
  } else {
    bitsize = bits_for_product(sizeint);
  }
  int smallidx = FIRSTIDX;
  while (smallidx < LASTIDX - 1 && magicints[smallidx] < mindiff) ++smallidx;
  put_be32(out, (uint32_t)smallidx);
  const int maxidx = std::min(LASTIDX - 1, smallidx + 8);
  const int minidx = maxidx - 8;
  int smaller = magicints[std::max(FIRSTIDX, smallidx - 1)] / 2;
  int smallnum = magicints[smallidx] / 2;
  unsigned sizesmall[3] = {(unsigned)magicints[smallidx], (unsigned)magicints[smallidx], (unsigned)magicints[smallidx]};
  const int larger = magicints[maxidx] / 2;
  BitWriter bits;
  int prevrun = -1;
  int prev[3] = {0, 0, 0};
  int64_t i = 0;
  auto near = [](const int* a, const int* b, int64_t limit) {
    return std::llabs((int64_t)a[0] - b[0]) < limit && std::llabs((int64_t)a[1] - b[1]) < limit &&
           std::llabs((int64_t)a[2] - b[2]) < limit;
  };
  while (i < natoms) {
    int* cur = &ints[3 * i];
    bool is_small = false;
    int is_smaller;
    if (smallidx < maxidx && i >= 1 && near(cur, prev, larger))
      is_smaller = 1;
    else if (smallidx > minidx)
      is_smaller = -1;
    else
      is_smaller = 0;
    if (i + 1 < natoms && near(cur, cur + 3, smallnum)) {
      for (int k = 0; k < 3; ++k) std::swap(cur[k], cur[3 + k]);
      is_small = true;
    }
    int tmp[3 * 8 + 3];
    for (int k = 0; k < 3; ++k) tmp[k] = cur[k] - minint[k];
    if (bitsize == 0) {
      for (int k = 0; k < 3; ++k) bits.put(bitsizeint[k], (unsigned)tmp[k]);
    } else {
      bits.put_ints(bitsize, sizeint, tmp);
    }
    for (int k = 0; k < 3; ++k) prev[k] = cur[k];
    cur += 3;
    ++i;
    int run = 0;
    if (!is_small && is_smaller == -1) is_smaller = 0;
    while (is_small && run < 8 * 3) {
      if (is_smaller == -1) {
        const int64_t dx = (int64_t)cur[0] - prev[0], dy = (int64_t)cur[1] - prev[1], dz = (int64_t)cur[2] - prev[2];
        if (dx * dx + dy * dy + dz * dz >= (int64_t)smaller * smaller) is_smaller = 0;
      }
      for (int k = 0; k < 3; ++k) {
        tmp[run++] = cur[k] - prev[k] + smallnum;
        prev[k] = cur[k];
      }
      ++i;
      cur += 3;
      is_small = i < natoms && near(cur, prev, smallnum);
    }
    if (run != prevrun || is_smaller != 0) {
      prevrun = run;
      bits.put(1, 1);
      bits.put(5, (unsigned)(run + is_smaller + 1));
    } else {
      bits.put(1, 0);
    }
    for (int k = 0; k < run; k += 3) bits.put_ints(smallidx, sizesmall, &tmp[k]);
    if (is_smaller != 0) {
      smallidx += is_smaller;
      if (is_smaller < 0) {
        smallnum = smaller;
        smaller = magicints[smallidx - 1] / 2;
      } else {
        smaller = smallnum;
        smallnum = magicints[smallidx] / 2;
      }
      sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)magicints[smallidx];
    }
  }
  put_be32(out, (uint32_t)bits.data.size());
  out.insert(out.end(), bits.data.begin(), bits.data.end());
  while (out.size() % 4) out.push_back(0);
  return true;
}

}  // namespace
}  // namespace cvp

using namespace cvp;

extern "C" {

int cvp_xtc_open(const char* path, cvp_xtc** out) {
  CVP_REQUIRE(path != nullptr && out != nullptr, "null argument");
  auto* f = new cvp_xtc();
  auto fail = [&](const std::string& why) {
    set_error(std::string(path) + ": " + why);
    if (f->map) munmap(const_cast<unsigned char*>(f->map), f->size);
    if (f->fd >= 0) close(f->fd);
    delete f;
    return CVP_ERR_INVALID;
  };
  f->fd = open(path, O_RDONLY);
  if (f->fd < 0) return fail("cannot open file");
  struct stat st;
  if (fstat(f->fd, &st) != 0 || st.st_size < 56) return fail("not an XTC file (too short)");
  f->size = (size_t)st.st_size;
  void* map = mmap(nullptr, f->size, PROT_READ, MAP_SHARED, f->fd, 0);
  if (map == MAP_FAILED) return fail("mmap failed");
  f->map = static_cast<const unsigned char*>(map);
  size_t pos = 0;
  while (pos + 56 <= f->size) {
    const unsigned char* p = f->map + pos;
    const int magic = (int32_t)be32(p);
    if (magic == 2023) return fail("XTC files with 64-bit block sizes (magic 2023) are not supported");
    if (magic != XTC_MAGIC) return fail("not an XTC file (bad magic number)");
    const int64_t natoms = (int32_t)be32(p + 4);
    if (natoms <= 0) return fail("corrupt XTC frame header");
    if (f->offsets.empty())
      f->num_atoms = natoms;
    else if (natoms != f->num_atoms)
      return fail("the number of atoms changes between frames");
    const int64_t block = coordinate_block_bytes(p + 52, f->size - pos - 52, natoms);
    if (block < 0 || pos + 52 + (size_t)block > f->size) break;  // truncated last frame: ignored
    if (natoms > 9 && f->precision == 0.0f) f->precision = be_float(p + 56);
    f->offsets.push_back((int64_t)pos);
    f->steps.push_back((int32_t)be32(p + 8));
    f->times.push_back(be_float(p + 12));
    pos += 52 + (size_t)block;
  }
  if (f->offsets.empty()) return fail("no complete frame in the file");
  *out = f;
  return CVP_OK;
}

int cvp_xtc_info(const cvp_xtc* f, int64_t* num_frames, int64_t* num_atoms, float* precision) {
  CVP_REQUIRE(f != nullptr, "null argument");
  if (num_frames) *num_frames = (int64_t)f->offsets.size();
  if (num_atoms) *num_atoms = f->num_atoms;
  if (precision) *precision = f->precision;
  return CVP_OK;
}

// Frames first .. first + count - 1 -> positions [count][N][3] float nm; box [count][9] doubles (rows
// a, b, c in nm), steps [count] and times [count] (ps) when not NULL.  `threads` <= 0: one per core.
int cvp_xtc_read(const cvp_xtc* f, int64_t first, int64_t count, float* positions, double* box,
                 int32_t* steps, float* times, int threads) {
  CVP_REQUIRE(f != nullptr && positions != nullptr, "null argument");
  CVP_REQUIRE(first >= 0 && count >= 0 && first + count <= (int64_t)f->offsets.size(), "frames outside the file");
  if (count == 0) return CVP_OK;
  int workers = threads > 0 ? threads : (int)std::max(1u, std::thread::hardware_concurrency());
  workers = (int)std::min<int64_t>(workers, count);
  const int64_t n = f->num_atoms;
  std::atomic<int64_t> next{0}, bad{-1};
  auto work = [&]() {
    for (;;) {
      const int64_t b = next.fetch_add(1);
      if (b >= count) return;
      const unsigned char* p = f->map + f->offsets[first + b];
      if (box)
        for (int k = 0; k < 9; ++k) box[9 * b + k] = (double)be_float(p + 16 + 4 * k);
      if (steps) steps[b] = f->steps[first + b];
      if (times) times[b] = f->times[first + b];
      if (!decode_coordinates(p + 52, n, positions + b * n * 3, nullptr)) bad.store(first + b);
    }
  };
  if (workers <= 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    pool.reserve(workers);
    for (int w = 0; w < workers; ++w) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  if (bad.load() >= 0) {
    set_error("corrupt XTC coordinate block in frame " + std::to_string(bad.load()));
    return CVP_ERR_INVALID;
  }
  return CVP_OK;
}

int cvp_xtc_close(cvp_xtc* f) {
  if (f == nullptr) return CVP_OK;
  if (f->map) munmap(const_cast<unsigned char*>(f->map), f->size);
  if (f->fd >= 0) close(f->fd);
  delete f;
  return CVP_OK;
}

// Appends (append != 0) or writes `count` frames: positions [count][N][3] float nm, box [count][9]
// doubles or NULL (zeros), steps / times [count] or NULL (0, 1, ... / 0.0).  precision <= 0: 1000.
int cvp_xtc_write(const char* path, int append, int64_t count, int64_t num_atoms, const float* positions,
                  const double* box, const int32_t* steps, const float* times, float precision) {
  CVP_REQUIRE(path != nullptr && (positions != nullptr || count == 0), "null argument");
  CVP_REQUIRE(count >= 0 && num_atoms > 0 && num_atoms <= INT_MAX, "invalid sizes");
  std::FILE* file = std::fopen(path, append ? "ab" : "wb");
  if (file == nullptr) {
    set_error(std::string(path) + ": cannot open file for writing");
    return CVP_ERR_INVALID;
  }
  std::vector<unsigned char> frame;
  for (int64_t b = 0; b < count; ++b) {
    frame.clear();
    put_be32(frame, (uint32_t)XTC_MAGIC);
    put_be32(frame, (uint32_t)num_atoms);
    put_be32(frame, (uint32_t)(steps ? steps[b] : (int32_t)b));
    put_be_float(frame, times ? times[b] : 0.0f);
    for (int k = 0; k < 9; ++k) put_be_float(frame, box ? (float)box[9 * b + k] : 0.0f);
    if (!encode_coordinates(positions + b * num_atoms * 3, num_atoms, precision, frame)) {
      std::fclose(file);
      set_error("coordinate too large for the requested XTC precision in frame " + std::to_string(b));
      return CVP_ERR_INVALID;
    }
    if (std::fwrite(frame.data(), 1, frame.size(), file) != frame.size()) {
      std::fclose(file);
      set_error(std::string(path) + ": write failed");
      return CVP_ERR_INVALID;
    }
  }
  std::fclose(file);
  return CVP_OK;
}

}  // extern "C"
