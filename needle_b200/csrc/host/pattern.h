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
};

constexpr uint32_t kBlobMagic = 0x424C444Eu;  // "NDLB"
constexpr int32_t kBlobVersion = 1;

// regex (UTF-16 code units) + flags -> CompiledPattern.  Throws SyntaxError / CompileError /
// TooLargeError / FlagsError (ast.h).  Restates DFACompiler.compileToBytes (DFACompiler.java:45-74).
CompiledPattern compile_pattern(const std::u16string& regex, int flags);

std::vector<uint8_t> serialize_pattern(const CompiledPattern& p);
// Throws std::runtime_error on a malformed blob.
CompiledPattern deserialize_pattern(const uint8_t* blob, size_t len);

}  // namespace ndl
