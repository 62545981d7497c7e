// CompiledPattern: everything the generated Matcher class of the reference carries as static state
// (DFAClassBuilder.java:57-128), as plain data, plus its binary serialisation (the "table blob").
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ndl {

enum TableId { kMatches = 0, kContainedIn = 1, kForwards = 2, kBackwards = 3 };

// One STATES_<SPEC> array + its wasAccepted<Spec> data (DFAClassBuilder.java:317-333, 701-728).
struct Table {
  int32_t n_states = 0;
  int32_t width = 1;                // 1: byte[] (<= 127 states), 2: short[] (DFAClassBuilder.java:79-85)
  int32_t max_char = 0;             // DFA.maxChar() of this DFA
  std::vector<uint8_t> accepting;   // n_states flags
  std::vector<int16_t> entries;     // n_states * stride, -1 = dead; entry[state * stride + byteClass]
  int n_accepting() const {
    int k = 0;
    for (uint8_t a : accepting) k += a;
    return k;
  }
};

enum ReverseMode { kReverseTable = 0, kReverseSingleChar = 1, kReverseFixedLength = 2 };

// The search accelerators a compiled needle class carries for indexForwards (DFAClassBuilder.java:365-429): PREFIX /
// SUFFIX / INFIX string constants, FIRST_BYTE_MASK, and the compile-time decisions of CompilationPolicy.java:44-57 and
// FindMethodSpec.java:63-89.  Result-preserving seeks; the GPU kernels do not use them (they read every byte), the CPU
// baseline does (oracle/needle_oracle.c, ndlo_index_forwards_accel).  Built by host/factorization.cpp.
struct Accel {
  bool present = false;  // false: a version-1 blob without this record
  bool use_prefix = false, use_suffix = false, use_infixes = false, use_max_start = false;
  bool can_seek_for_predicate = false;   // FindMethodSpec.canSeekForPredicate
  bool has_first_byte_mask = false;      // dfaSearch.initialAsciiBytes().isPresent()
  bool byte_check_first_char = false;    // FindMethodSpec.doByteCheckForFirstCharacter
  bool post_prefix_accepting = false, follow_accepting = false;
  bool inner_must_call_was_accepted = false;  // DFAClassBuilder.isInnerLoopMustCallWasAccepted
  int32_t post_prefix_state = 0;         // dfaSearch.after(prefix)
  int32_t follow_state = 0;              // dfaSearch.forwardFollowingState()
  int32_t pred_kind = 0;                 // generatePredicate: 0 none, 1 c == a, 2 a <= c <= b, 3 c == a || c == b
  int32_t pred_a = 0, pred_b = 0;
  std::u16string prefix, suffix, infix;  // empty when unused
  uint8_t first_byte_mask[129] = {};     // index min(c, 128)
};

struct CompiledPattern {
  int32_t flags = 0;
  int32_t min_length = 0;
  int32_t max_length = -1;
  int32_t stride = 0;            // getEffectiveByteClassCount(byteClassCount)
  int32_t byte_class_count = 0;  // ByteClasses.byteClassCount
  int32_t reverse_mode = kReverseTable;
  int32_t reverse_char = 0;
  // BYTE_CLASSES[0..65535] (DFAClassBuilder.java:269-305).  Stored as unsigned class ids; the
  // reference's ids are Java bytes, so only patterns with <= 127 ids behave meaningfully there
  //  - see compile.cpp for what happens above that.
  std::vector<uint16_t> class_map;  // 65536 entries
  Table tables[4];
  Accel accel;
};

constexpr uint32_t kBlobMagic = 0x424C444Eu;  // "NDLB"
constexpr int32_t kBlobVersion = 2;  // 2: + accelerator record (version-1 blobs are still read)

// regex (UTF-16 code units) + flags -> CompiledPattern.  Throws SyntaxError / CompileError /
// TooLargeError / FlagsError (ast.h).  Restates DFACompiler.compileToBytes (DFACompiler.java:45-74).
CompiledPattern compile_pattern(const std::u16string& regex, int flags);

std::vector<uint8_t> serialize_pattern(const CompiledPattern& p);
// Throws std::runtime_error on a malformed blob.
CompiledPattern deserialize_pattern(const uint8_t* blob, size_t len);

}  // namespace ndl
