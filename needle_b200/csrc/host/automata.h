// NFA program, subset construction, minimisation and byte classes of the host-side compiler.
// Restates RegexInstrBuilder.java, NFA.java:225-266, NFAToDFACompiler.java, StateSet.java, DFA.java,
// MinimizeDFA.java, CharRange.java:81-157 of the reference (paths under
// needle-compiler/src/main/java/com/justinblank/strings/).
#pragma once
#include <cstdint>
#include <deque>
#include <memory>
#include <utility>
#include <vector>

#include "ast.h"

namespace ndl {

enum class Op { CharRange, Jump, Split, Match };

struct Instr {
  Op op = Op::Match;
  uint16_t start = 'a', end = 'a';
  int jump = -1;
  std::vector<int> split;
  int priority = 0;
  bool null = false;  // placeholder slot (`instrs.add(null)`)
};

// RegexInstrBuilder.createNFA(node, leftMostLongest) (RegexInstrBuilder.java:24-35)
std::vector<Instr> build_program(const Node* ast, bool leftmost_longest);

enum class ConversionMode { Basic, ContainedIn, DfaSearch };

struct Dfa;
struct DfaState {
  bool accepting = false;
  int number = 0;
  std::vector<std::pair<CharRange, DfaState*>> transitions;
  // DFA.addTransition (DFA.java:63-83)
  void add_transition(CharRange r, DfaState* target);
  DfaState* step(uint16_t c) const;
  int char_total() const;
};

struct Dfa {
  std::deque<DfaState> arena;
  std::vector<DfaState*> states;  // states[0] is the root
  DfaState* root() const { return states[0]; }
  int count() const { return static_cast<int>(states.size()); }
  DfaState* new_state(bool accepting, int number) {
    arena.emplace_back();
    DfaState* s = &arena.back();
    s->accepting = accepting;
    s->number = number;
    states.push_back(s);
    return s;
  }
  int max_char() const;              // DFA.maxChar (DFA.java:384-398)
  void prune_dead_states();          // DFA.pruneDeadStates (DFA.java:745-792)
  DfaState* after(const std::u16string& s) const;
};

// NFAToDFACompiler._compile only (no pruning / minimisation): used by structure tests
// (NFAToDFACompilerTest.java:11-35 counts these states).
std::unique_ptr<Dfa> subset_construction(const std::vector<Instr>& prog, ConversionMode mode);
// MinimizeDFA.minimizeDFA (MinimizeDFA.java:18-38)
std::unique_ptr<Dfa> minimize(const Dfa& dfa);
// NFAToDFACompiler.compile (NFAToDFACompiler.java:24-32): construct, prune dead states, minimise.
std::unique_ptr<Dfa> compile_dfa(const std::vector<Instr>& prog, ConversionMode mode);

struct ByteClasses {
  bool present = false;          // Optional.empty() when there are >= 255 range groups
  std::vector<uint8_t> ranges;   // 65537 entries, Java `byte` bit patterns (ids >= 128 are negative there)
  int catch_all = 0;
  int byte_class_count = 0;
  // Not in the reference: the same class ids without the Java-byte truncation and with U+FFFF given
  // its real class; used for patterns the reference's byte[] scheme cannot represent (> 127 ids).
  int n_groups = 0;
  std::vector<uint16_t> wide;    // 65536 entries, ids 1..n_groups, 0 = in no transition range
};
// DFA.byteClasses (DFA.java:438-463)
ByteClasses byte_classes(const Dfa& dfa);

}  // namespace ndl
