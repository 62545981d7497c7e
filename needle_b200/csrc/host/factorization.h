// Literal factors of a regex and the search accelerators the reference derives from them.
//
// Restates needle-compiler/src/main/java/com/justinblank/strings/Factorization.java (prefix / suffix / factor / required
// factor algebra over the AST, after Navarro & Raffinot 5.5.2; each AST node's bestFactors(): RegexAST/*.java) and
// CompilationPolicy.java:44-75 (which of them indexForwards uses), plus the decisions DFAClassBuilder.createIndexMethod
// takes from the search DFA (DFAClassBuilder.java:365-429; DFA.java:358-366, 669-742; FindMethodSpec.java:63-89).
//
// None of this can change a match result: the accelerators only skip positions from which the search automaton would
// sit in its root state.  The GPU kernels do not use them (they read every byte at HBM speed); they are carried in the
// blob because they are part of what a compiled needle class holds as static state (PREFIX / SUFFIX / INFIX /
// FIRST_BYTE_MASK constants), and because the CPU baseline the GPU path is compared with must have them - they are what
// makes the JVM version fast on sparse inputs.
#pragma once
#include <cstdint>
#include <optional>
#include <set>
#include <string>

#include "ast.h"
#include "automata.h"
#include "pattern.h"

namespace ndl {

// CompilationPolicy constants (CompilationPolicy.java:12-20)
constexpr int kPredicateRangeSizeCutoff = 6;
constexpr int kFactorizationMaxCharRangeSize = 4;
constexpr int kFactorizationMaxRepetitionCount = 2;
constexpr int kThresholdForCalculatingMaxStart = 4;
constexpr double kMaxFrequencyForInitialCharCheck = .12;

struct Factorization {
  using Set = std::set<std::u16string>;
  using OptSet = std::optional<Set>;  // nullopt = Java null ("cannot be computed")
  OptSet all, suffixes, prefixes, factors, required;
  int min_length = 0, max_length = -1;

  static Factorization of_string(const std::u16string& s);  // Factorization(String)
  static Factorization of_range(uint16_t start, uint16_t end);
  static Factorization empty();
  void unite(const Factorization& o);        // union
  void concatenate(const Factorization& o);
  Factorization counted_repetition(int min, int max);  // (mutates *this like the reference when min != 0)
  std::optional<std::u16string> shared_prefix() const;
  std::optional<std::u16string> shared_suffix() const;
};

// Node.bestFactors() + Factorization.buildFactorization (Factorization.java:101-106)
Factorization build_factorization(const Node* node);

// CompilationPolicy.create + the search-DFA decisions of createIndexMethod -> the blob's accelerator record.
Accel build_accel(const Factorization& f, const Dfa& search);

}  // namespace ndl
