// C ABI, host half: regex -> blob, blob introspection, error plumbing.  See include/needle_b200.h.
#include <cstdlib>
#include <cstring>
#include <string>

#include "capi_internal.h"
#include "host/ast.h"
#include "host/pattern.h"
#include "host_chunks.h"
#include "host_staging.h"
#include "needle_b200.h"

namespace ndl {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

// UTF-8 -> UTF-16 code units (BMP and surrogate pairs for the rest, like a Java string literal).
static bool utf8_to_utf16(const char* s, std::u16string& out) {
  const unsigned char* p = reinterpret_cast<const unsigned char*>(s);
  while (*p) {
    uint32_t cp;
    int extra;
    if (*p < 0x80) { cp = *p; extra = 0; }
    else if ((*p & 0xE0) == 0xC0) { cp = *p & 0x1F; extra = 1; }
    else if ((*p & 0xF0) == 0xE0) { cp = *p & 0x0F; extra = 2; }
    else if ((*p & 0xF8) == 0xF0) { cp = *p & 0x07; extra = 3; }
    else return false;
    p++;
    for (int i = 0; i < extra; i++, p++) {
      if ((*p & 0xC0) != 0x80) return false;
      cp = (cp << 6) | (*p & 0x3F);
    }
    if (cp >= 0x10000) {
      cp -= 0x10000;
      out.push_back(static_cast<char16_t>(0xD800 + (cp >> 10)));
      out.push_back(static_cast<char16_t>(0xDC00 + (cp & 0x3FF)));
    } else {
      out.push_back(static_cast<char16_t>(cp));
    }
  }
  return true;
}

static int compile_to_blob(const std::u16string& regex, int flags, uint8_t** blob_out, size_t* blob_len_out) {
  if (!blob_out || !blob_len_out) return fail(NDL_EINVAL, "blob_out / blob_len_out must not be NULL");
  *blob_out = nullptr;
  *blob_len_out = 0;
  try {
    CompiledPattern p = compile_pattern(regex, flags);
    std::vector<uint8_t> bytes = serialize_pattern(p);
    uint8_t* mem = static_cast<uint8_t*>(std::malloc(bytes.size()));
    if (!mem) return fail(NDL_ENOMEM, "out of memory");
    std::memcpy(mem, bytes.data(), bytes.size());
    *blob_out = mem;
    *blob_len_out = bytes.size();
    return NDL_OK;
  } catch (const SyntaxError& e) {
    return fail(NDL_ESYNTAX, e.what());
  } catch (const TooLargeError& e) {
    return fail(NDL_ETOOLARGE, e.what());
  } catch (const FlagsError& e) {
    return fail(NDL_EFLAGS, e.what());
  } catch (const CompileError& e) {
    return fail(NDL_ECOMPILE, e.what());
  } catch (const std::bad_alloc&) {
    return fail(NDL_ENOMEM, "out of memory");
  } catch (const std::exception& e) {
    return fail(NDL_ECOMPILE, e.what());
  }
}

}  // namespace ndl

using namespace ndl;

extern "C" {

int ndl_compile(const uint16_t* regex_utf16, size_t n_chars, int flags, uint8_t** blob_out, size_t* blob_len_out) {
  if (!regex_utf16 && n_chars) return fail(NDL_EINVAL, "regex string cannot be null");
  std::u16string regex(reinterpret_cast<const char16_t*>(regex_utf16), n_chars);
  return compile_to_blob(regex, flags, blob_out, blob_len_out);
}

int ndl_compile_utf8(const char* regex_utf8, int flags, uint8_t** blob_out, size_t* blob_len_out) {
  if (!regex_utf8) return fail(NDL_EINVAL, "regex string cannot be null");
  std::u16string regex;
  if (!utf8_to_utf16(regex_utf8, regex)) return fail(NDL_EINVAL, "regex is not valid UTF-8");
  return compile_to_blob(regex, flags, blob_out, blob_len_out);
}

void ndl_blob_free(uint8_t* blob) { std::free(blob); }

int ndl_blob_info_get(const uint8_t* blob, size_t blob_len, ndl_blob_info* out) {
  if (!out) return fail(NDL_EINVAL, "out must not be NULL");
  try {
    CompiledPattern p = deserialize_pattern(blob, blob_len);
    out->version = kBlobVersion;
    out->flags = p.flags;
    out->min_length = p.min_length;
    out->max_length = p.max_length;
    out->stride = p.stride;
    out->byte_class_count = p.byte_class_count;
    out->reverse_mode = p.reverse_mode;
    out->reverse_char = p.reverse_char;
    for (int k = 0; k < 4; k++) {
      out->n_states[k] = p.tables[k].n_states;
      out->entry_width[k] = p.tables[k].width;
      out->max_char[k] = p.tables[k].max_char;
      out->n_accepting[k] = p.tables[k].n_accepting();
    }
    return NDL_OK;
  } catch (const std::exception& e) {
    return fail(NDL_EBLOB, e.what());
  }
}

const char* ndl_last_error(void) { return g_last_error.c_str(); }

const char* ndl_version(void) { return "needle_b200 0.2.0 (blob v2, sm_100a)"; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// Test hook (not part of include/needle_b200.h): exposes intermediate results of the host pipeline so
// that the reference's structure tests (NFAToDFACompilerTest.java:11-35, DFATest.java:186-352) can be
// replayed.  conversion_mode: 0 BASIC, 1 CONTAINED_IN, 2 DFA_SEARCH.  The NFA is built with
// leftMostLongest = true, like RegexInstrBuilder.createNFA(ast) which those tests go through.
// ---------------------------------------------------------------------------------------------
#include "host/automata.h"

extern "C" int ndl_debug_dfa(const uint16_t* regex_utf16, size_t n_chars, int flags, int conversion_mode,
                             int32_t* raw_states, int32_t* minimized_states, uint8_t* byte_classes_65536) {
  try {
    std::u16string regex(reinterpret_cast<const char16_t*>(regex_utf16), n_chars);
    Ast ast;
    Node* node = parse_regex(ast, regex, flags);
    std::vector<Instr> prog = build_program(node, true);
    ConversionMode mode = conversion_mode == 0 ? ConversionMode::Basic
                          : conversion_mode == 1 ? ConversionMode::ContainedIn : ConversionMode::DfaSearch;
    std::unique_ptr<Dfa> raw = subset_construction(prog, mode);
    if (raw_states) *raw_states = raw->count();
    raw->prune_dead_states();
    std::unique_ptr<Dfa> m = minimize(*raw);
    if (minimized_states) *minimized_states = m->count();
    if (byte_classes_65536) {
      ByteClasses bc = byte_classes(*m);
      if (!bc.present) return fail(NDL_ECOMPILE, "no byte classes");
      for (int c = 0; c < 65536; c++) byte_classes_65536[c] = bc.ranges[c];
    }
    return NDL_OK;
  } catch (const SyntaxError& e) {
    return fail(NDL_ESYNTAX, e.what());
  } catch (const std::exception& e) {
    return fail(NDL_ECOMPILE, e.what());
  }
}

// Test hook (not part of include/needle_b200.h; host only): how the NDL_MEM_HOST path of ndl_match_batch would cut a batch
// into pipeline chunks (host_chunks.h - the same functions capi_device.cu calls).  bounds receives the chunk boundaries
// (bounds[0] = 0, bounds[k + 1] = end of chunk k; room for max_chunks + 1 entries), flags per chunk bit 0 = offsets equally
// spaced, bit 1 = offsets non-decreasing.  Returns the number of chunks.
extern "C" int ndl_debug_plan_chunks(const uint64_t* offsets, uint64_t line_chars, uint64_t n, int char_width, uint64_t chunk_bytes,
                                     int max_chunks, uint64_t* bounds, int32_t* flags) {
  const uint64_t total = offsets ? offsets[n] - offsets[0] : n * line_chars;
  const int n_chunks = host_chunk_count(static_cast<size_t>(total) * char_width, n, chunk_bytes, max_chunks);
  int used = 0;
  uint64_t i0 = 0;
  bounds[0] = 0;
  for (int k = 0; k < n_chunks && i0 < n; k++) {
    const uint64_t i1 = host_chunk_end(offsets, n, i0, k, n_chunks);
    OffsetScan sc{true, true, line_chars};
    if (offsets) sc = scan_offsets(offsets + i0, i1 - i0);
    flags[used] = (sc.uniform ? 1 : 0) | (sc.monotonic ? 2 : 0);
    bounds[++used] = i1;
    i0 = i1;
  }
  return used;
}

// Test hook (not part of include/needle_b200.h; host only): `callers` threads copy disjoint slices of src to dst through the copy
// pool AT THE SAME TIME - what the replica threads of a multi-device pattern do with pageable input.  Returns the pool's threads.
extern "C" int ndl_debug_parallel_copy(uint8_t* dst, const uint8_t* src, uint64_t bytes, int callers) {
  if (callers < 1) callers = 1;
  std::vector<std::thread> threads;
  const uint64_t slice = (bytes + callers - 1) / callers;
  for (int k = 0; k < callers; k++) {
    const uint64_t lo = std::min(bytes, slice * k), hi = std::min(bytes, lo + slice);
    threads.emplace_back([=] {
      for (uint64_t off = lo; off < hi; off += 5u << 20)  // ring-sized pieces, as h2d_copy issues them
        ndl::CopyPool::instance().copy(dst + off, src + off, static_cast<size_t>(std::min<uint64_t>(5u << 20, hi - off)));
    });
  }
  for (auto& t : threads) t.join();
  return ndl::CopyPool::instance().threads();
}

// Test hook (not part of include/needle_b200.h): the byte classes EACH of the four automata of a pattern would have on
// its own (DFA.byteClasses applied to it), in table order MATCHES, CONTAINEDIN, FORWARDS, BACKWARDS; out = 4 x 65536 ids.
// The reference derives ONE class map, from the search automaton only, and shares it between all four tables
// (DFAClassBuilder.java:67-76, with the TODO "test whether it matters that the four DFAs can have different
// byteClasses"): where another automaton distinguishes chars that the search automaton does not, that automaton's
// table is coarser than the automaton.  The generative differential test uses this to tell such patterns.
extern "C" int ndl_debug_class_maps(const uint16_t* regex_utf16, size_t n_chars, int flags, uint16_t* out) {
  try {
    std::u16string regex(reinterpret_cast<const char16_t*>(regex_utf16), n_chars);
    Ast ast;
    Node* node = parse_regex(ast, regex, flags);
    const bool lml = (flags & NDL_LEFTMOST_LONGEST) == NDL_LEFTMOST_LONGEST;
    std::vector<Instr> forward = build_program(node, lml);
    std::vector<Instr> reversed = build_program(ast.reversed(node), lml);
    std::unique_ptr<Dfa> dfas[4] = {compile_dfa(forward, ConversionMode::Basic), compile_dfa(forward, ConversionMode::ContainedIn),
                                    compile_dfa(forward, ConversionMode::DfaSearch), compile_dfa(reversed, ConversionMode::Basic)};
    for (int k = 0; k < 4; k++) {
      ByteClasses bc = byte_classes(*dfas[k]);
      for (int c = 0; c < 65536; c++) out[k * 65536 + c] = bc.wide[c];
    }
    return NDL_OK;
  } catch (const SyntaxError& e) {
    return fail(NDL_ESYNTAX, e.what());
  } catch (const std::exception& e) {
    return fail(NDL_ECOMPILE, e.what());
  }
}
