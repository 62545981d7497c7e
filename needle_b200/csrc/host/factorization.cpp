// See factorization.h.  Restates Factorization.java, the bestFactors() of each RegexAST node, CompilationPolicy.create
// and the search-DFA decisions of DFAClassBuilder.createIndexMethod.
#include "factorization.h"

#include <algorithm>

#include "jhash.h"

namespace ndl {

namespace {

using Set = Factorization::Set;
using OptSet = Factorization::OptSet;

// Factorization.concatenateStrings (:214-233): null on either side -> EMPTY set; an empty side -> that side.
Set concatenate_strings(const OptSet& a, const OptSet& b) {
  if (!a || !b) return Set();
  if (a->empty()) return *a;
  if (b->empty()) return *b;
  Set out;
  for (const auto& x : *a)
    for (const auto& y : *b) out.insert(x + y);
  return out;
}

// Factorization.best (:243-275): a smaller set of longer factors is preferred.  May return null.
OptSet best(const OptSet& s1, const OptSet& s2) {
  if (!s1 || s1->empty()) return s2;
  if (!s2 || s2->empty()) return s1;
  auto min_len = [](const Set& s) { size_t m = SIZE_MAX; for (auto& x : s) m = std::min(m, x.size()); return m; };
  auto max_len = [](const Set& s) { size_t m = 0; for (auto& x : s) m = std::max(m, x.size()); return m; };
  auto sum_len = [](const Set& s) { size_t m = 0; for (auto& x : s) m += x.size(); return m; };
  if (s1->size() > s2->size()) {
    if (min_len(*s1) <= max_len(*s2)) return s2;
  } else if (s2->size() > s1->size()) {
    if (min_len(*s2) <= max_len(*s1)) return s1;
  }
  if (sum_len(*s2) > sum_len(*s1)) return s2;
  return s1;
}

OptSet set_union(const OptSet& a, const OptSet& b) {
  if (!a || !b) return std::nullopt;
  Set out = *a;
  out.insert(b->begin(), b->end());
  return out;
}

std::optional<std::u16string> shared_prefix_of(const OptSet& prefixes) {
  // Factorization.getSharedPrefix(Set) (:318-345)
  if (!prefixes || prefixes->empty()) return std::nullopt;
  if (prefixes->size() == 1 && prefixes->begin()->empty()) return std::nullopt;
  bool first = true;
  std::u16string shared;
  for (const auto& p : *prefixes) {
    if (first) {
      shared = p;
      first = false;
      continue;
    }
    size_t i = 0;
    while (i < shared.size() && i < p.size() && shared[i] == p[i]) i++;
    if (i == 0) return std::nullopt;
    shared.resize(i);
  }
  return shared;
}

Factorization best_factors(const Node* n) {
  if (!n) throw std::logic_error("null node");  // NullPointerException in the reference (e.g. `a|`): a compilation failure
  switch (n->kind) {
    case NodeKind::Literal: return Factorization::of_string(n->lit);                      // LiteralNode.java:37-39
    case NodeKind::Range: return Factorization::of_range(n->range.start, n->range.end);   // CharRangeNode.java:37-39
    case NodeKind::Union: {                                                               // Union.java:38-43
      Factorization l = best_factors(n->a);
      l.unite(best_factors(n->b));
      return l;
    }
    case NodeKind::Concat: {                                                              // Concatenation.java:51-56
      Factorization l = best_factors(n->a);
      l.concatenate(best_factors(n->b));
      return l;
    }
    case NodeKind::Repetition: return Factorization::empty();                             // Repetition.java:31-33
    case NodeKind::Counted: {                                                             // CountedRepetition.java:37-39
      Factorization f = best_factors(n->a);
      return f.counted_repetition(n->min, n->max);
    }
    case NodeKind::LParen: break;
  }
  throw std::logic_error("bestFactors of a parenthesis node");  // UnsupportedOperationException
}

struct JStringHash {  // java.lang.String.hashCode
  uint32_t operator()(const std::u16string& s) const {
    uint32_t h = 0;
    for (char16_t c : s) h = 31u * h + c;
    return h;
  }
};

}  // namespace

Factorization Factorization::of_string(const std::u16string& s) {
  Factorization f;
  f.all = Set{s};
  f.suffixes = Set{s};
  f.prefixes = Set{s};
  f.factors = Set{s};
  f.required = Set();
  if (!s.empty()) f.required->insert(s);
  return f;
}

Factorization Factorization::of_range(uint16_t start, uint16_t end) {
  // Factorization.fromRange (:81-94)
  if (static_cast<int>(end) - static_cast<int>(start) > kFactorizationMaxCharRangeSize) return empty();
  if (start == end) return of_string(std::u16string(1, static_cast<char16_t>(start)));
  Set strings;
  for (int c = start; c <= end; c++) strings.insert(std::u16string(1, static_cast<char16_t>(c)));
  Factorization f;
  f.all = f.suffixes = f.prefixes = f.factors = strings;
  f.required = Set();
  return f;
}

Factorization Factorization::empty() {
  Factorization f;  // all four sets null, requiredFactors an empty set (:100-102)
  f.required = Set();
  return f;
}

void Factorization::unite(const Factorization& o) {
  // Factorization.union (:143-188): set union per component, null wins; required factors are INTERSECTED
  all = set_union(all, o.all);
  prefixes = set_union(prefixes, o.prefixes);
  suffixes = set_union(suffixes, o.suffixes);
  factors = set_union(factors, o.factors);
  if (!required || !o.required) {
    required = std::nullopt;
  } else {
    Set keep;
    for (const auto& x : *required)
      if (o.required->count(x)) keep.insert(x);
    required = keep;
  }
}

void Factorization::concatenate(const Factorization& o) {
  // Factorization.concatenate (:190-212)
  const OptSet local_suffixes = suffixes, local_prefixes = prefixes, local_factors = factors, local_all = all;
  if (!local_all || !o.all) all = std::nullopt;
  else all = concatenate_strings(local_all, o.all);
  prefixes = best(local_prefixes, OptSet(concatenate_strings(local_all, o.prefixes)));
  suffixes = best(o.suffixes, OptSet(concatenate_strings(local_suffixes, o.all)));
  const OptSet first_factors = best(local_factors, o.factors);
  factors = best(first_factors, OptSet(concatenate_strings(local_suffixes, o.prefixes)));
  Set req = required ? *required : Set();
  if (o.required) req.insert(o.required->begin(), o.required->end());
  required = req;
}

Factorization Factorization::counted_repetition(int min, int max) {
  // Factorization.countedRepetition (:281-296).  With min != 0 the reference accumulates into `this` - and builds every
  // repetition from the already-updated `this`; reproduced literally.
  if (max > kFactorizationMaxRepetitionCount) return empty();
  Factorization fresh;
  Factorization* acc = this;
  if (min == 0) {
    fresh = of_string(std::u16string());
    acc = &fresh;
  }
  for (int i = min; i <= max; i++) {
    Factorization rep = of_string(std::u16string());
    for (int j = 0; j < i; j++) rep.concatenate(*this);
    acc->unite(rep);
  }
  return *acc;
}

std::optional<std::u16string> Factorization::shared_prefix() const { return shared_prefix_of(prefixes); }

std::optional<std::u16string> Factorization::shared_suffix() const {
  // getSharedSuffix (:347-353): shared prefix of the reversed suffixes, reversed
  if (!suffixes || suffixes->empty()) return std::nullopt;
  Set rev;
  for (auto s : *suffixes) {
    std::reverse(s.begin(), s.end());
    rev.insert(s);
  }
  auto p = shared_prefix_of(OptSet(rev));
  if (p) std::reverse(p->begin(), p->end());
  return p;
}

Factorization build_factorization(const Node* node) {
  Factorization f = best_factors(node);
  f.min_length = Ast::min_length(node);
  f.max_length = Ast::max_length(node);
  return f;
}

Accel build_accel(const Factorization& f, const Dfa& search) {
  static const double kDefaultWeights[128] = {
#include "char_distribution.inc"
  };
  Accel a;
  a.present = true;
  // --- CompilationPolicy.create (:44-57)
  const auto sp = f.shared_prefix(), ss = f.shared_suffix();
  const bool has_max = f.max_length != kNoMax;
  a.use_prefix = sp && !sp->empty();
  a.use_suffix = ss && has_max && !(sp && *sp == *ss);
  // getRequiredInfixes (:133-138) in java.util.HashSet iteration order, chooseInfix (:63-75): the first longest one
  std::u16string infix;
  bool any_infix = false;
  if (f.required) {
    auto infixes = JHashSet<std::u16string, JStringHash>::withExpected(f.required->size());
    for (const auto& s : *f.required) infixes.add(s);
    if (sp) infixes.remove(*sp);
    if (ss) infixes.remove(*ss);
    for (const auto& s : infixes.items()) {
      if (!any_infix || s.size() > infix.size()) infix = s;
      any_infix = true;
    }
  }
  a.use_infixes = any_infix && has_max;
  a.use_max_start = f.min_length > kThresholdForCalculatingMaxStart;
  if (a.use_prefix) a.prefix = *sp;
  if (a.use_suffix) a.suffix = *ss;
  if (a.use_infixes) a.infix = infix;

  // --- the search DFA's root (createIndexMethod, DFAClassBuilder.java:365-429)
  const DfaState* root = search.root();
  std::vector<std::pair<CharRange, DfaState*>> leaving;  // DFA.getTransitionsLeavingZeroState (:722-730)
  for (const auto& t : root->transitions)
    if (t.second->number != 0) leaving.push_back(t);
  if (a.use_prefix) {
    const DfaState* after = search.after(a.prefix);
    if (!after) throw CompileError("No DFA state available after consuming prefix. This should be impossible");
    a.post_prefix_state = after->number;
    a.post_prefix_accepting = after->accepting;
  }
  bool same_target = true;  // allForwardTransitionsLeadToSameState (:358-366)
  for (size_t i = 0; i + 1 < leaving.size(); i++) same_target = same_target && leaving[i].second == leaving[i + 1].second;
  bool is_predicate = false;  // forwardTransitionIsPredicate (:689-703)
  if (leaving.size() == 1) {
    is_predicate = 1 + (leaving[0].first.end - leaving[0].first.start) <= kPredicateRangeSizeCutoff;
  } else if (leaving.size() == 2 && leaving[0].first.single() && leaving[1].first.single()) {
    is_predicate = std::abs(static_cast<int>(leaving[0].first.start) - static_cast<int>(leaving[1].first.start)) == 32;
  }
  a.can_seek_for_predicate = is_predicate && same_target && !root->accepting;
  if (a.can_seek_for_predicate) {
    a.follow_state = leaving[0].second->number;
    a.follow_accepting = leaving[0].second->accepting;
    if (leaving.size() == 1) {  // generatePredicate (:482-505)
      a.pred_kind = leaving[0].first.single() ? 1 : 2;
      a.pred_a = leaving[0].first.start;
      a.pred_b = leaving[0].first.end;
    } else {
      a.pred_kind = 3;
      a.pred_a = leaving[0].first.start;
      a.pred_b = leaving[1].first.end;
    }
  }
  // initialAsciiBytes (:706-726)
  a.has_first_byte_mask = true;
  for (const auto& t : leaving) {
    if (t.first.end < 128) {
      for (int c = t.first.start; c <= t.first.end; c++) a.first_byte_mask[c] = 1;
    } else if (t.first.start == 128 && t.first.end == 0xFFFF) {
      a.first_byte_mask[128] = 1;
    } else {
      a.has_first_byte_mask = false;
      break;
    }
  }
  if (!a.has_first_byte_mask) std::fill(a.first_byte_mask, a.first_byte_mask + 129, 0);
  // doByteCheckForFirstCharacter (FindMethodSpec.java:79-89).  (weight() would index weights[128] for the catch-all
  // entry and throw in the reference; that entry counts as weight 0 here.)
  a.byte_check_first_char = false;
  if (!root->accepting && a.has_first_byte_mask) {
    double w = 0;
    for (int c = 0; c < 128; c++)
      if (a.first_byte_mask[c]) w += kDefaultWeights[c];
    a.byte_check_first_char = w < kMaxFrequencyForInitialCharCheck &&
                              !(a.use_prefix || a.use_infixes || a.use_suffix || a.can_seek_for_predicate);
  }
  // isInnerLoopMustCallWasAccepted (:473-480)
  a.inner_must_call_was_accepted = root->accepting || (a.use_prefix && a.post_prefix_accepting);
  return a;
}

}  // namespace ndl
