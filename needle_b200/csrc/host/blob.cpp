// Binary "table blob": the serialised CompiledPattern.
//
// Analogue of the class file Precompile.precompile writes (precompile/Precompile.java:30-53) and of the
// BYTE_CLASS_STRING_* wire format the generated <clinit> decodes (ByteClassUtil.java:8-12, 50-120): it
// is what ships to disk, across JNI, and to the other GPUs.  Little-endian, versioned, checksummed.
//
//   u32 magic "NDLB" | i32 version | i32 flags | i32 min_length | i32 max_length | i32 stride |
//   i32 byte_class_count | i32 reverse_mode | i32 reverse_char |
//   u32 n_runs | n_runs x { u16 last_char, u16 class }      -- BYTE_CLASSES as runs (DFAClassBuilder.java:280-299)
//   4 x { i32 n_states | i32 width | i32 max_char | u8 accepting[n_states] (padded to 4) |
//         i16 entries[n_states * stride] (padded to 4) }    -- MATCHES, CONTAINEDIN, FORWARDS, BACKWARDS
//   version >= 2: the accelerator record (pattern.h Accel; DFAClassBuilder.java:365-429)
//     u32 "ACCL" | u32 flag bits | i32 post_prefix_state | i32 follow_state | i32 pred_kind | i32 pred_a | i32 pred_b |
//     3 x { u32 n | u16 chars[n] (padded to 4) }  -- PREFIX, SUFFIX, INFIX | u8 first_byte_mask[129] (padded to 4)
//   u32 fnv1a of everything before it
#include <cstring>
#include <stdexcept>

#include "pattern.h"

namespace ndl {

namespace {

struct Writer {
  std::vector<uint8_t> buf;
  void u32(uint32_t v) {
    for (int i = 0; i < 4; i++) buf.push_back(static_cast<uint8_t>(v >> (8 * i)));
  }
  void i32(int32_t v) { u32(static_cast<uint32_t>(v)); }
  void u16(uint16_t v) {
    buf.push_back(static_cast<uint8_t>(v));
    buf.push_back(static_cast<uint8_t>(v >> 8));
  }
  void pad4() {
    while (buf.size() % 4) buf.push_back(0);
  }
};

struct Reader {
  const uint8_t* p;
  size_t len, pos = 0;
  void need(size_t n) const {
    if (pos + n > len) throw std::runtime_error("truncated pattern blob");
  }
  uint32_t u32() {
    need(4);
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) v |= static_cast<uint32_t>(p[pos + i]) << (8 * i);
    pos += 4;
    return v;
  }
  int32_t i32() { return static_cast<int32_t>(u32()); }
  uint16_t u16() {
    need(2);
    uint16_t v = static_cast<uint16_t>(p[pos] | (p[pos + 1] << 8));
    pos += 2;
    return v;
  }
  void pad4() {
    while (pos % 4) {
      need(1);
      pos++;
    }
  }
};

constexpr uint32_t kAccelMagic = 0x4C434341u;  // "ACCL"

uint32_t fnv1a(const uint8_t* p, size_t n) {
  uint32_t h = 2166136261u;
  for (size_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 16777619u;
  }
  return h;
}

}  // namespace

std::vector<uint8_t> serialize_pattern(const CompiledPattern& p) {
  Writer w;
  w.u32(kBlobMagic);
  w.i32(kBlobVersion);
  w.i32(p.flags);
  w.i32(p.min_length);
  w.i32(p.max_length);
  w.i32(p.stride);
  w.i32(p.byte_class_count);
  w.i32(p.reverse_mode);
  w.i32(p.reverse_char);
  std::vector<std::pair<uint16_t, uint16_t>> runs;
  for (int c = 0; c < 65536; c++) {
    if (c == 65535 || p.class_map[c + 1] != p.class_map[c]) runs.push_back({static_cast<uint16_t>(c), p.class_map[c]});
  }
  w.u32(static_cast<uint32_t>(runs.size()));
  for (auto& r : runs) {
    w.u16(r.first);
    w.u16(r.second);
  }
  for (int k = 0; k < 4; k++) {
    const Table& t = p.tables[k];
    w.i32(t.n_states);
    w.i32(t.width);
    w.i32(t.max_char);
    for (uint8_t a : t.accepting) w.buf.push_back(a);
    w.pad4();
    for (int16_t e : t.entries) w.u16(static_cast<uint16_t>(e));
    w.pad4();
  }
  const Accel& a = p.accel;
  w.u32(kAccelMagic);
  w.u32((a.use_prefix ? 1u : 0) | (a.use_suffix ? 2u : 0) | (a.use_infixes ? 4u : 0) | (a.use_max_start ? 8u : 0) |
        (a.can_seek_for_predicate ? 16u : 0) | (a.has_first_byte_mask ? 32u : 0) | (a.byte_check_first_char ? 64u : 0) |
        (a.post_prefix_accepting ? 128u : 0) | (a.follow_accepting ? 256u : 0) | (a.inner_must_call_was_accepted ? 512u : 0) |
        (a.present ? 1024u : 0));
  w.i32(a.post_prefix_state);
  w.i32(a.follow_state);
  w.i32(a.pred_kind);
  w.i32(a.pred_a);
  w.i32(a.pred_b);
  for (const std::u16string* str : {&a.prefix, &a.suffix, &a.infix}) {
    w.u32(static_cast<uint32_t>(str->size()));
    for (char16_t c : *str) w.u16(static_cast<uint16_t>(c));
    w.pad4();
  }
  for (int i = 0; i < 129; i++) w.buf.push_back(a.first_byte_mask[i]);
  w.pad4();
  w.u32(fnv1a(w.buf.data(), w.buf.size()));
  return w.buf;
}

CompiledPattern deserialize_pattern(const uint8_t* blob, size_t len) {
  if (!blob || len < 44) throw std::runtime_error("pattern blob too short");
  Reader r{blob, len};
  if (r.u32() != kBlobMagic) throw std::runtime_error("not a needle_b200 pattern blob (bad magic)");
  int32_t version = r.i32();
  if (version != 1 && version != kBlobVersion) throw std::runtime_error("unsupported pattern blob version " + std::to_string(version));
  uint32_t stored = 0;
  std::memcpy(&stored, blob + len - 4, 4);  // little-endian hosts only (x86-64 / aarch64)
  if (fnv1a(blob, len - 4) != stored) throw std::runtime_error("pattern blob checksum mismatch");
  CompiledPattern p;
  p.flags = r.i32();
  p.min_length = r.i32();
  p.max_length = r.i32();
  p.stride = r.i32();
  p.byte_class_count = r.i32();
  p.reverse_mode = r.i32();
  p.reverse_char = r.i32();
  if (p.stride <= 0 || p.stride > 65536 || p.byte_class_count > p.stride || p.reverse_mode < 0 || p.reverse_mode > 2)
    throw std::runtime_error("pattern blob header out of range");
  uint32_t n_runs = r.u32();
  if (n_runs == 0 || n_runs > 65536) throw std::runtime_error("pattern blob class map out of range");
  p.class_map.assign(65536, 0);
  int c = 0;
  for (uint32_t i = 0; i < n_runs; i++) {
    int last = r.u16();
    int cls = r.u16();
    if (last < c || cls >= p.stride) throw std::runtime_error("pattern blob class map malformed");
    for (; c <= last; c++) p.class_map[c] = static_cast<uint16_t>(cls);
  }
  if (c != 65536) throw std::runtime_error("pattern blob class map does not cover all chars");
  for (int k = 0; k < 4; k++) {
    Table& t = p.tables[k];
    t.n_states = r.i32();
    t.width = r.i32();
    t.max_char = r.i32();
    if (t.n_states <= 0 || t.n_states > 16383 || (t.width != 1 && t.width != 2) || t.max_char < 0 || t.max_char > 0xFFFF)
      throw std::runtime_error("pattern blob table header out of range");
    r.need(t.n_states);
    t.accepting.assign(blob + r.pos, blob + r.pos + t.n_states);
    r.pos += t.n_states;
    r.pad4();
    size_t n = static_cast<size_t>(t.n_states) * p.stride;
    t.entries.resize(n);
    for (size_t i = 0; i < n; i++) {
      int16_t e = static_cast<int16_t>(r.u16());
      if (e < -1 || e >= t.n_states) throw std::runtime_error("pattern blob transition out of range");
      t.entries[i] = e;
    }
    r.pad4();
  }
  if (version >= 2) {
    if (r.u32() != kAccelMagic) throw std::runtime_error("pattern blob accelerator record missing");
    Accel& a = p.accel;
    const uint32_t f = r.u32();
    a.use_prefix = f & 1; a.use_suffix = f & 2; a.use_infixes = f & 4; a.use_max_start = f & 8;
    a.can_seek_for_predicate = f & 16; a.has_first_byte_mask = f & 32; a.byte_check_first_char = f & 64;
    a.post_prefix_accepting = f & 128; a.follow_accepting = f & 256; a.inner_must_call_was_accepted = f & 512;
    a.present = f & 1024;
    a.post_prefix_state = r.i32();
    a.follow_state = r.i32();
    a.pred_kind = r.i32();
    a.pred_a = r.i32();
    a.pred_b = r.i32();
    const int n_fwd = p.tables[kForwards].n_states;
    if (a.post_prefix_state < 0 || a.post_prefix_state >= n_fwd || a.follow_state < 0 || a.follow_state >= n_fwd || a.pred_kind < 0 ||
        a.pred_kind > 3)
      throw std::runtime_error("pattern blob accelerator record out of range");
    for (std::u16string* str : {&a.prefix, &a.suffix, &a.infix}) {
      const uint32_t n = r.u32();
      if (n > 65536) throw std::runtime_error("pattern blob accelerator string too long");
      str->clear();
      for (uint32_t i = 0; i < n; i++) str->push_back(static_cast<char16_t>(r.u16()));
      r.pad4();
    }
    r.need(129);
    for (int i = 0; i < 129; i++) a.first_byte_mask[i] = blob[r.pos + i];
    r.pos += 129;
    r.pad4();
  }
  if (r.pos + 4 != len) throw std::runtime_error("pattern blob has trailing bytes");
  return p;
}

}  // namespace ndl
