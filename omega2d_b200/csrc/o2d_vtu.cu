// o2d_vtu.cu -- the .vtu file image of a particle collection, encoded on the GPU.
//
// Replaces Points<S>::write_vtk (reference src/Points.h:556-740) with its helpers
// VtkXmlWriter::unpackArray / writeDataArray (src/VtkXmlWriter.h:71-106; base64 through
// cppcodec::base64_rfc4648, reference lib/cppcodec).  The file is ~55 bytes per particle, all but ~1.3 kB
// of it base64 text: position and velocity as interleaved (x, y, 0) float triples, the cell arrays
// (connectivity = 0..n-1, offsets = 1..n, types = 1) and one Float32 array per scalar.  The reference
// interleaves into a temporary, encodes into a std::string and pushes the text through tinyxml2's
// character-escaping printer -- 0.37 s per million particles on this container's host.
//
// Here ONE kernel launch encodes every array section straight from the SoA arrays (which already live on
// the GPU between time steps): each thread turns 12 input bytes into 16 characters and stores them with
// one 16-byte store; the interleave, the iota arrays and the constant types array are generated in
// registers and never exist in memory.  The host writes the XML skeleton (a few hundred bytes between
// the sections) and copies each section to its place in the image.  HBM-bound: 24 B read + 54.7 B written
// per particle with strengths and radii.
#include "o2d_context.h"

#include <cstring>
#include <string>

using namespace o2d;

namespace {

constexpr int kVtuThreads = 256;
constexpr int kMaxSections = 9;

enum : int { SEC_XY0 = 0, SEC_WORDS = 1, SEC_IOTA = 2, SEC_ONES = 3 };

struct VtuSection {
  int kind;
  int start;                   // SEC_IOTA: first value
  const uint32_t* a;           // SEC_XY0: x ; SEC_WORDS: the array
  const uint32_t* b;           // SEC_XY0: y
  char* out;                   // 16-byte aligned staging area of this section
  unsigned long long nbytes;   // raw bytes to encode
  unsigned int first_block;    // blocks [first_block, next section's first_block) work on this section
};
struct VtuArgs {
  VtuSection sec[kMaxSections];
  int nsec;
  int n;
};

// four 6-bit values (one per byte) -> four characters of the RFC 4648 alphabet:
//   0..25 'A'+i | 26..51 'a'+(i-26) | 52..61 '0'+(i-52) | 62 '+' | 63 '/'
// = i + 65 + 6 [i>=26] - 75 [i>=52] - 15 [i>=62] + 3 [i>=63], evaluated for the four bytes at once: every
// byte of the result is a valid character (< 256), so the packed integer sum is exact whatever
// borrows happen on the way.
__device__ __forceinline__ uint32_t b64_chars(uint32_t idx4) {
  const uint32_t ge26 = ((idx4 + 0x66666666u) >> 7) & 0x01010101u;   // i + 102 >= 128
  const uint32_t ge52 = ((idx4 + 0x4C4C4C4Cu) >> 7) & 0x01010101u;   // i + 76
  const uint32_t ge62 = ((idx4 + 0x42424242u) >> 7) & 0x01010101u;   // i + 66
  const uint32_t ge63 = ((idx4 + 0x41414141u) >> 7) & 0x01010101u;   // i + 65
  return idx4 + 0x41414141u + 6u * ge26 + 3u * ge63 - 75u * ge52 - 15u * ge62;
}

// v = b0<<16 | b1<<8 | b2 (three stream bytes) -> the four characters, first character in the low byte
__device__ __forceinline__ uint32_t b64_group(uint32_t v) {
  const uint32_t idx4 = (v >> 18) | ((v >> 4) & 0x3F00u) | ((v << 10) & 0x3F0000u) | ((v << 24) & 0x3F000000u);
  return b64_chars(idx4);
}

__global__ void __launch_bounds__(kVtuThreads)
vtu_encode_kernel(const VtuArgs a) {
  // which section does this block work on (at most 9 entries, block-uniform)
  int k = 0;
#pragma unroll 1
  for (int j = 1; j < a.nsec; ++j)
    if (blockIdx.x >= a.sec[j].first_block) k = j;
  const VtuSection& s = a.sec[k];
  const unsigned long long t = (unsigned long long)(blockIdx.x - s.first_block) * kVtuThreads + threadIdx.x;
  const unsigned long long byte0 = 12ull * t;
  if (byte0 >= s.nbytes) return;
  const unsigned long long left = s.nbytes - byte0;
  const int valid = left >= 12ull ? 12 : (int)left;

  uint32_t w0 = 0, w1 = 0, w2 = 0;
  if (s.kind == SEC_XY0) {                     // unpackArray: (x, y, 0) per point, src/VtkXmlWriter.h:71-87
    w0 = s.a[t];
    w1 = s.b[t];
  } else if (s.kind == SEC_WORDS) {
    const unsigned long long i = 3ull * t;
    w0 = s.a[i];
    if (valid > 4) w1 = s.a[i + 1];
    if (valid > 8) w2 = s.a[i + 2];
  } else if (s.kind == SEC_IOTA) {             // std::iota, src/Points.h:618-632
    const uint32_t i = (uint32_t)(3ull * t) + (uint32_t)s.start;
    w0 = i;
    if (valid > 4) w1 = i + 1;
    if (valid > 8) w2 = i + 2;
  } else {                                     // std::fill(v, 1), src/Points.h:640-641
    w0 = valid >= 4 ? 0x01010101u : (0x01010101u >> (8 * (4 - valid)));
    w1 = valid >= 8 ? 0x01010101u : (valid > 4 ? (0x01010101u >> (8 * (8 - valid))) : 0u);
    w2 = valid >= 12 ? 0x01010101u : (valid > 8 ? (0x01010101u >> (8 * (12 - valid))) : 0u);
  }

  uint4 o;
  o.x = b64_group(__byte_perm(w0, 0u, 0x4012));                       // bytes 0 1 2
  o.y = b64_group(__byte_perm(w0, w1, 0x3345) & 0x00FFFFFFu);         // bytes 3 4 5
  o.z = b64_group(__byte_perm(w1, w2, 0x0234) & 0x00FFFFFFu);         // bytes 6 7 8
  o.w = b64_group(__byte_perm(w2, 0u, 0x4123));                       // bytes 9 10 11
  if (valid < 12 && (valid % 3) != 0) {
    // the array ends inside a 3-byte group: '=' padding (only the last thread of a section gets here)
    const int g = valid / 3;                   // group holding the last 1 or 2 bytes
    const uint32_t pad = (valid % 3) == 1 ? 0x3D3D0000u : 0x3D000000u;
    const uint32_t keep = (valid % 3) == 1 ? 0x0000FFFFu : 0x00FFFFFFu;
    if (g == 0) o.x = (o.x & keep) | pad;
    else if (g == 1) o.y = (o.y & keep) | pad;
    else if (g == 2) o.z = (o.z & keep) | pad;
    else o.w = (o.w & keep) | pad;
  }
  *reinterpret_cast<uint4*>(s.out + 16ull * t) = o;
}

// ---- host side: the XML skeleton ------------------------------------------------------------------------
const char kB64[] = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";

size_t b64_len(size_t nbytes) { return 4 * ((nbytes + 2) / 3); }

void b64_host(std::string& o, const unsigned char* p, size_t n) {
  size_t i = 0;
  for (; i + 3 <= n; i += 3) {
    o += kB64[p[i] >> 2];
    o += kB64[((p[i] & 3) << 4) | (p[i + 1] >> 4)];
    o += kB64[((p[i + 1] & 15) << 2) | (p[i + 2] >> 6)];
    o += kB64[p[i + 2] & 63];
  }
  if (n - i == 1) {
    o += kB64[p[i] >> 2]; o += kB64[(p[i] & 3) << 4]; o += "==";
  } else if (n - i == 2) {
    o += kB64[p[i] >> 2]; o += kB64[((p[i] & 3) << 4) | (p[i + 1] >> 4)]; o += kB64[(p[i + 1] & 15) << 2]; o += '=';
  }
}

// One array section of the file: the text in front of its base64 payload (ends with the encoded UInt32
// length header) and what the kernel encodes.
struct Piece {
  std::string text;      // skeleton text preceding the payload
  int kind = -1;         // -1: no payload (trailer)
  const float* a = nullptr;
  const float* b = nullptr;
  int start = 0;
  size_t nbytes = 0;
};

// VtkXmlWriter::writeDataArray (src/VtkXmlWriter.h:94-106): format="binary", a blank, the base64 of the UInt32
// length OF THE ENCODED TEXT (:99-100), the data, a blank; closeElement() then closes on the same line.
void open_array(std::string& t, const char* indent, const char* attribs, size_t nbytes) {
  t += indent; t += "<DataArray "; t += attribs; t += " format=\"binary\"> ";
  const uint32_t enc = (uint32_t)b64_len(nbytes);
  b64_host(t, reinterpret_cast<const unsigned char*>(&enc), 4);
}

// the whole file as alternating text / payload pieces; layout follows src/Points.h:556-740 line by line
int build_pieces(int n, const float* x, const float* y, const float* s, const float* r, const float* w,
                 const float* e, const float* u, const float* v, double time, Piece* pc) {
  const char* i2 = "        ";
  const char* i3 = "            ";
  const char* i4 = "                ";
  const std::string close = " </DataArray>\n";
  const size_t nn = (size_t)n;
  int k = 0;
  std::string t;
  // VtkXmlWriter constructor (src/VtkXmlWriter.h:40-58), FieldData/TimeValue (src/Points.h:575-588)
  t = "<?xml version=\"1.0\"?>\n"
      "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt32\">\n"
      "    <UnstructuredGrid>\n";
  t += i2; t += "<FieldData>\n";
  open_array(t, i3, "Name=\"TimeValue\" NumberOfTuples=\"1\" type=\"Float64\"", sizeof(double));
  b64_host(t, reinterpret_cast<const unsigned char*>(&time), sizeof(double));
  t += close;
  t += i2; t += "</FieldData>\n";
  t += i2; t += "<Piece NumberOfCells=\"" + std::to_string(n) + "\" NumberOfPoints=\"" + std::to_string(n) + "\">\n";
  t += i3; t += "<Points>\n";
  open_array(t, i4, "Name=\"position\" NumberOfComponents=\"3\" type=\"Float32\"", 12 * nn);
  pc[k].text = t; pc[k].kind = SEC_XY0; pc[k].a = x; pc[k].b = y; pc[k].nbytes = 12 * nn; ++k;
  // Cells (src/Points.h:609-652)
  t = close; t += i3; t += "</Points>\n"; t += i3; t += "<Cells>\n";
  open_array(t, i4, "Name=\"connectivity\" type=\"Int32\"", 4 * nn);
  pc[k].text = t; pc[k].kind = SEC_IOTA; pc[k].start = 0; pc[k].nbytes = 4 * nn; ++k;
  t = close;
  open_array(t, i4, "Name=\"offsets\" type=\"Int32\"", 4 * nn);
  pc[k].text = t; pc[k].kind = SEC_IOTA; pc[k].start = 1; pc[k].nbytes = 4 * nn; ++k;
  t = close;
  open_array(t, i4, "Name=\"types\" type=\"UInt8\"", nn);
  pc[k].text = t; pc[k].kind = SEC_ONES; pc[k].nbytes = nn; ++k;
  // PointData (src/Points.h:654-709): the Scalars list names circulation, radius, vorticity only
  t = close; t += i3; t += "</Cells>\n";
  std::string scal;
  if (s) scal += "circulation,";
  if (r) scal += "radius,";
  if (w) scal += "vorticity,";
  t += i3;
  if (scal.size() > 1) { scal.pop_back(); t += "<PointData Scalars=\"" + scal + "\" Vectors=\"velocity\">\n"; }
  else t += "<PointData Vectors=\"velocity\">\n";
  const struct { const float* p; const char* attribs; } scalars[4] = {
      {s, "Name=\"circulation\" type=\"Float32\""}, {r, "Name=\"radius\" type=\"Float32\""},
      {w, "Name=\"vorticity\" type=\"Float32\""}, {e, "Name=\"shearrate\" type=\"Float32\""}};
  bool first = true;
  for (const auto& sc : scalars) {
    if (!sc.p) continue;
    if (!first) t = close;
    first = false;
    open_array(t, i4, sc.attribs, 4 * nn);
    pc[k].text = t; pc[k].kind = SEC_WORDS; pc[k].a = sc.p; pc[k].nbytes = 4 * nn; ++k;
  }
  if (!first) t = close;
  open_array(t, i4, "Name=\"velocity\" NumberOfComponents=\"3\" type=\"Float32\"", 12 * nn);
  pc[k].text = t; pc[k].kind = SEC_XY0; pc[k].a = u; pc[k].b = v; pc[k].nbytes = 12 * nn; ++k;
  // PointData, Piece, and finish() (src/VtkXmlWriter.h:128-134)
  t = close; t += i3; t += "</PointData>\n"; t += i2; t += "</Piece>\n";
  t += "    </UnstructuredGrid>\n</VTKFile>\n";
  pc[k].text = t; pc[k].kind = -1; ++k;
  return k;
}

size_t image_bytes(const Piece* pc, int npc) {
  size_t tot = 0;
  for (int k = 0; k < npc; ++k) tot += pc[k].text.size() + (pc[k].kind >= 0 ? b64_len(pc[k].nbytes) : 0);
  return tot;
}

inline size_t align256(size_t b) { return (b + 255) / 256 * 256; }

}  // namespace

// the UInt32 length header wraps at 4 GiB of encoded text per array (16 n characters for the vectors)
static int check_count(int n) {
  if (n <= 0) return fail(O2D_ERR_INVALID, "write_vtk needs n > 0 (the reference asserts it, src/Points.h:557)");
  if (n >= (1 << 28)) return fail(O2D_ERR_UNSUPPORTED, "n >= 2^28: the UInt32 array header of the format overflows");
  return O2D_OK;
}

extern "C" int o2d_vtu_points_size(int n, int has_s, int has_r, int has_w, int has_e, size_t* bytes) {
  if (int rc = check_count(n)) return rc;
  if (!bytes) return fail(O2D_ERR_INVALID, "bytes is NULL");
  Piece pc[kMaxSections + 1];
  const float* one = reinterpret_cast<const float*>(sizeof(float));   // only tested against NULL
  const int npc = build_pieces(n, one, one, has_s ? one : nullptr, has_r ? one : nullptr, has_w ? one : nullptr,
                               has_e ? one : nullptr, one, one, 0.0, pc);
  *bytes = image_bytes(pc, npc);
  return O2D_OK;
}

extern "C" int o2d_vtu_points(o2d_context* c, int mem, int n, const float* x, const float* y,
                              const float* s, const float* r, const float* w, const float* e,
                              const float* u, const float* v, double time,
                              char* out, size_t cap, size_t* written) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (int rc = check_count(n)) return rc;
  if (!x || !y || !u || !v) return fail(O2D_ERR_INVALID, "x, y, u, v must not be NULL");
  if (!out) return fail(O2D_ERR_INVALID, "out is NULL");

  // ---- inputs on the device ---------------------------------------------------------------------------
  const float* in[8] = {x, y, s, r, w, e, u, v};
  if (mem == O2D_MEM_HOST) {
    int cnt = 0;
    for (const float* p : in) cnt += p != nullptr;
    CU(c->hsrc.reserve((size_t)cnt * padded_floats(n)));
    Stager S(c, &c->hsrc);
    for (auto& p : in) {
      if (!p) continue;
      float* d = S.take(n);
      if (int rc = o2d_h2d(c, d, p, n)) return rc;
      p = d;
    }
  }

  Piece pc[kMaxSections + 1];
  const int npc = build_pieces(n, in[0], in[1], in[2], in[3], in[4], in[5], in[6], in[7], time, pc);
  const size_t total = image_bytes(pc, npc);
  if (written) *written = total;
  if (total > cap) return fail(O2D_ERR_INVALID, "output buffer too small: %zu bytes needed, %zu given", total, cap);

  // ---- one launch encodes every section into the staging buffer ---------------------------------------------
  size_t stage = 0;
  for (int k = 0; k < npc; ++k)
    if (pc[k].kind >= 0) stage += align256(b64_len(pc[k].nbytes) + 16);
  CU(c->vtu.reserve(stage));
  VtuArgs a;
  a.nsec = 0; a.n = n;
  unsigned int nblocks = 0;
  size_t off = 0;
  for (int k = 0; k < npc; ++k) {
    if (pc[k].kind < 0) continue;
    VtuSection& sec = a.sec[a.nsec++];
    sec.kind = pc[k].kind; sec.start = pc[k].start;
    sec.a = reinterpret_cast<const uint32_t*>(pc[k].a);
    sec.b = reinterpret_cast<const uint32_t*>(pc[k].b);
    sec.out = c->vtu.as<char>() + off;
    sec.nbytes = pc[k].nbytes;
    sec.first_block = nblocks;
    const size_t groups = (pc[k].nbytes + 11) / 12;
    nblocks += (unsigned int)((groups + kVtuThreads - 1) / kVtuThreads);
    off += align256(b64_len(pc[k].nbytes) + 16);
  }
  {
    PairTimer timer(c);
    vtu_encode_kernel<<<nblocks, kVtuThreads, 0, c->stream>>>(a);
    CU(cudaGetLastError());
    c->launches += 1;
  }

  // ---- assemble the image: skeleton text from the host, payloads from the device -----------------------------
  size_t pos = 0;
  int isec = 0;
  for (int k = 0; k < npc; ++k) {
    std::memcpy(out + pos, pc[k].text.data(), pc[k].text.size());
    pos += pc[k].text.size();
    if (pc[k].kind < 0) continue;
    const size_t enc = b64_len(pc[k].nbytes);
    CU(cudaMemcpyAsync(out + pos, a.sec[isec].out, enc, cudaMemcpyDeviceToHost, c->stream));
    pos += enc;
    ++isec;
  }
  CU(cudaStreamSynchronize(c->stream));
  return O2D_OK;
}

extern "C" int o2d_write_vtu_points(o2d_context* c, int mem, const char* path, int n, const float* x, const float* y,
                                    const float* s, const float* r, const float* w, const float* e,
                                    const float* u, const float* v, double time) {
  if (int rc = o2d_check_common(c, mem)) return rc;
  if (!path) return fail(O2D_ERR_INVALID, "path is NULL");
  size_t need = 0;
  if (int rc = o2d_vtu_points_size(n, s != nullptr, r != nullptr, w != nullptr, e != nullptr, &need)) return rc;
  // grow-only pinned image buffer: the payload copies run at full PCIe rate
  if (need > c->vtu_host_cap) {
    if (c->vtu_host) { CU(cudaFreeHost(c->vtu_host)); c->vtu_host = nullptr; c->vtu_host_cap = 0; }
    const size_t want = need + need / 4 + 4096;
    CU(cudaHostAlloc(&c->vtu_host, want, cudaHostAllocDefault));
    c->vtu_host_cap = want;
  }
  size_t wrote = 0;
  if (int rc = o2d_vtu_points(c, mem, n, x, y, s, r, w, e, u, v, time, static_cast<char*>(c->vtu_host), c->vtu_host_cap, &wrote))
    return rc;
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return fail(O2D_ERR_INVALID, "cannot open %s for writing", path);
  const size_t put = std::fwrite(c->vtu_host, 1, wrote, fp);
  const int bad = std::fclose(fp);
  if (put != wrote || bad) return fail(O2D_ERR_INVALID, "short write to %s", path);
  return O2D_OK;
}
