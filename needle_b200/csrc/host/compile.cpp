// regex -> CompiledPattern: the orchestration the reference does in DFACompiler.compileToBytes
// (DFACompiler.java:45-74) and the table layout decisions of DFAClassBuilder (stride rounding :240-253,
// byte-vs-short tables :79-85, shared byte classes :67-76, transition strings DFAStateTransitions.java:30-62
// decoded by ByteClassUtil.java:50-120, accepting arrays :701-728, reverse-scan choice :588-614, :119-121).
// The output is data for the GPU instead of JVM bytecode.
#include <algorithm>

#include "ast.h"
#include "automata.h"
#include "factorization.h"
#include "needle_b200.h"
#include "pattern.h"

namespace ndl {

namespace {

// DFAClassBuilder.getEffectiveByteClassCount (:240-253)
int effective_byte_class_count(int n) {
  if (n > 16) return n;  // THRESHOLD_TO_ROUND_UP_ALL_BYTECLASSES
  if (n < 3) return n;
  if (n < 4) return 4;
  if (n < 8) return 8;
  if (n < 16) return 16;
  return n;
}

// Table in the reference's own representation: what ByteClassUtil.fillMultipleByteClassesFromString*
// leaves in STATES_<SPEC> after decoding DFAStateTransitions.buildByteClassString for every state.
// First transition to claim a (state, class) cell wins (`bytesWitnessed`), classes are looked up per
// char so a transition range spanning several classes fills several cells, and because
// BYTE_CLASSES[0xFFFF] is never filled (DFA.java:451) any range ending at U+FFFF also claims class 0.
Table reference_table(const Dfa& dfa, const ByteClasses& bc, int stride) {
  Table t;
  t.n_states = dfa.count();
  t.width = dfa.count() > 127 ? 2 : 1;
  t.max_char = dfa.max_char();
  t.accepting.assign(t.n_states, 0);
  t.entries.assign(static_cast<size_t>(t.n_states) * stride, -1);
  // next index at which the class id changes, to skip over runs
  std::vector<int> run_end(65536);
  run_end[65535] = 65535;
  for (int i = 65534; i >= 0; i--) run_end[i] = (bc.ranges[i] == bc.ranges[i + 1]) ? run_end[i + 1] : i;
  for (const DfaState* s : dfa.states) {
    t.accepting[s->number] = s->accepting ? 1 : 0;
    std::vector<char> witnessed(256, 0);
    for (const auto& tr : s->transitions) {
      int i = tr.first.start;
      while (i <= tr.first.end) {
        int8_t cls = static_cast<int8_t>(bc.ranges[i]);
        if (cls >= 0 && !witnessed[cls]) {  // `byteClass == 0 || byteClass > largestSeenByteClass(=0)`: negative ids are skipped
          witnessed[cls] = 1;
          if (cls < stride) t.entries[static_cast<size_t>(s->number) * stride + cls] = static_cast<int16_t>(tr.second->number);
        }
        i = run_end[i] + 1;
      }
    }
  }
  return t;
}

// Exact table for patterns outside the reference's byte-class scheme: cell = the DFA's own transition.
Table exact_table(const Dfa& dfa, const ByteClasses& bc, int stride) {
  Table t;
  t.n_states = dfa.count();
  t.width = dfa.count() > 127 ? 2 : 1;
  t.max_char = 0xFFFF;
  t.accepting.assign(t.n_states, 0);
  t.entries.assign(static_cast<size_t>(t.n_states) * stride, -1);
  std::vector<int> rep(stride, -1);  // a representative char per class
  for (int c = 65535; c >= 0; c--) rep[bc.wide[c]] = c;
  for (const DfaState* s : dfa.states) {
    t.accepting[s->number] = s->accepting ? 1 : 0;
    for (int cls = 0; cls < stride; cls++) {
      if (rep[cls] < 0) continue;
      const DfaState* nx = s->step(static_cast<uint16_t>(rep[cls]));
      if (nx) t.entries[static_cast<size_t>(s->number) * stride + cls] = static_cast<int16_t>(nx->number);
    }
  }
  return t;
}

// Byte classes come from the search DFA only and are shared by all four tables
// (DFAClassBuilder.java:67-76); a char in no range of the search DFA has class 0.  For the exact
// scheme the classes must separate chars for all four DFAs, so they are computed on their union.
ByteClasses union_byte_classes(const Dfa* const dfas[4]) {
  Dfa merged;
  int base = 0;
  for (int k = 0; k < 4; k++) {
    std::vector<DfaState*> mine;
    for (const DfaState* s : dfas[k]->states) mine.push_back(merged.new_state(s->accepting, base + s->number));
    for (const DfaState* s : dfas[k]->states)
      for (const auto& tr : s->transitions) mine[s->number]->transitions.push_back({tr.first, mine[tr.second->number]});
    base += dfas[k]->count();
  }
  return byte_classes(merged);
}

}  // namespace

CompiledPattern compile_pattern(const std::u16string& regex, int flags) {
  if ((flags & ~NDL_ALL_FLAGS) != 0) throw FlagsError("Unrecognized flags=" + std::to_string(flags));
  try {
    Ast ast;
    Node* node = parse_regex(ast, regex, flags);
    CompiledPattern p;
    p.flags = flags;
    // Factorization.buildFactorization (Factorization.java:101-106): only min/max length reach the results; the literal
    // factors feed the search accelerators (p.accel, below)
    const Factorization factorization = build_factorization(node);
    p.min_length = factorization.min_length;
    p.max_length = factorization.max_length;

    const bool lml = (flags & NDL_LEFTMOST_LONGEST) == NDL_LEFTMOST_LONGEST;
    std::vector<Instr> forward = build_program(node, lml);
    std::vector<Instr> reversed = build_program(ast.reversed(node), lml);

    std::unique_ptr<Dfa> dfa = compile_dfa(forward, ConversionMode::Basic);
    std::unique_ptr<Dfa> contained = compile_dfa(forward, ConversionMode::ContainedIn);
    std::unique_ptr<Dfa> dfa_reversed = compile_dfa(reversed, ConversionMode::Basic);
    std::unique_ptr<Dfa> search = compile_dfa(forward, ConversionMode::DfaSearch);

    // DFACompiler.checkForOverLongDFAs (:76-83)
    const Dfa* all[4] = {dfa.get(), contained.get(), search.get(), dfa_reversed.get()};  // TableId order
    for (const Dfa* d : all)
      if (d->count() > 32767 / 2) throw TooLargeError("Can't compile DFAs with more than 16383 states");

    ByteClasses bc = byte_classes(*search);
    if (bc.present && bc.byte_class_count <= 128) {
      p.byte_class_count = bc.byte_class_count;
      p.stride = effective_byte_class_count(bc.byte_class_count);
      p.class_map.assign(65536, 0);
      for (int c = 0; c < 65536; c++) p.class_map[c] = bc.ranges[c];
      for (int k = 0; k < 4; k++) p.tables[k] = reference_table(*all[k], bc, p.stride);
    } else {
      // More class ids than a Java byte can hold.  The reference either mis-indexes its tables (128..254
      // ids) or drops to per-state range comparisons (>= 255 ids, DFAClassBuilder.java:796-842)
      // whose meaning is the DFA itself; both cases get exact tables over unsigned class ids here.
      ByteClasses ubc = union_byte_classes(all);
      p.byte_class_count = ubc.n_groups + 1;
      p.stride = ubc.n_groups + 1;
      p.class_map = ubc.wide;
      for (int k = 0; k < 4; k++) p.tables[k] = exact_table(*all[k], ubc, p.stride);
    }

    p.accel = build_accel(factorization, *search);

    // How find() derives start() from end() (DFAClassBuilder.java:119-121, 588-614, 640-657)
    if (p.max_length != kNoMax && p.min_length == p.max_length) {
      p.reverse_mode = kReverseFixedLength;  // Factorization.canOnlyHaveOneLength
    } else {
      p.reverse_mode = kReverseTable;
      // generateSingleCharacterReverseScan: the *matching* DFA's root has exactly one single-char
      // transition (DFA.firstStateCharacters :368-382 ignores multi-char ranges) and no other state can
      // consume that char (DFA.hasNonPrefix :817-828).
      if (dfa->count() != 1 && p.min_length != 0) {
        std::vector<uint16_t> cs;
        for (const auto& tr : dfa->root()->transitions)
          if (tr.first.single() && std::find(cs.begin(), cs.end(), tr.first.start) == cs.end()) cs.push_back(tr.first.start);
        if (cs.size() == 1) {
          bool non_prefix = false;
          for (const DfaState* s : dfa->states)
            if (s != dfa->root() && s->step(cs[0])) non_prefix = true;
          if (!non_prefix) {
            p.reverse_mode = kReverseSingleChar;
            p.reverse_char = cs[0];
          }
        }
      }
    }
    return p;
  } catch (const SyntaxError&) {
    throw;
  } catch (const TooLargeError&) {
    throw;
  } catch (const std::exception& e) {
    throw CompileError(std::string("Failed to create pattern for regex: ") + e.what());
  }
}

}  // namespace ndl
