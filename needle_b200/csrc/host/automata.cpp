// Host-side automata pipeline: AST -> Thompson program -> DFA (3 conversion modes) -> pruned ->
// minimised -> byte classes.  This stays on the host (BASELINE.json north_star); only its output - the
// transition tables - goes to the GPU.
//
// Every function restates one method of the reference; the citation is on the function.  Where the
// reference iterates a java.util.HashSet<Integer> and the order can decide a tie, JIntSet (jhash.h)
// reproduces that order.
#include "automata.h"

#include <algorithm>
#include <climits>
#include <map>
#include <set>
#include <unordered_map>

#include "jhash.h"

namespace ndl {

// ------------------------------------------------------------------------------------------------
// RegexInstrBuilder
// ------------------------------------------------------------------------------------------------
namespace {

Instr mk_jump(int target, int priority) {
  Instr i;
  i.op = Op::Jump;
  i.jump = target;
  i.priority = std::max(0, priority);
  return i;
}
Instr mk_split(std::vector<int> targets, int priority) {
  for (int t : targets)
    if (t < 0) throw std::invalid_argument("split target cannot be < 0");
  Instr i;
  i.op = Op::Split;
  i.split = std::move(targets);
  i.priority = std::max(0, priority);
  return i;
}
Instr mk_range(uint16_t s, uint16_t e, int priority) {
  Instr i;
  i.op = Op::CharRange;
  i.start = s;
  i.end = e;
  i.priority = std::max(0, priority);
  return i;
}
Instr mk_null() {
  Instr i;
  i.null = true;
  return i;
}

class ProgramBuilder {
 public:
  explicit ProgramBuilder(bool lml) : leftmost_longest_(lml) {}

  std::vector<Instr> build(const Node* ast) {
    create_partial(ast);
    Instr m;
    m.op = Op::Match;
    m.priority = std::max(0, priority_for_match());
    instrs_.push_back(m);
    resolve_jumps();
    return instrs_;
  }

 private:
  bool leftmost_longest_;
  int max_priority_ = 1;  // STARTING_PRIORITY
  std::vector<Instr> instrs_;

  const Instr& at(int i) const {
    if (i < 0 || i >= static_cast<int>(instrs_.size())) throw std::out_of_range("instr index");
    if (instrs_[i].null) throw std::logic_error("null instr");
    return instrs_[i];
  }

  // RegexInstrBuilder.priorityForMatch (:39-58)
  int priority_for_match() const {
    int match_index = static_cast<int>(instrs_.size());
    int priority = INT_MAX;
    for (const Instr& in : instrs_) {
      if (in.null) throw std::logic_error("null instr");
      if (in.op == Op::Jump) {
        if (in.jump == match_index) priority = std::min(priority, in.priority);
      } else if (in.op == Op::Split) {
        for (int t : in.split)
          if (t == match_index) {
            priority = std::min(priority, in.priority);
            break;
          }
      }
    }
    return priority;
  }

  // RegexInstrBuilder.getResolvedJump (:83-91)
  int resolved_jump(int jump) const {
    const Instr* target = &at(jump);
    int resolved = -1;
    while (target->op == Op::Jump) {
      resolved = target->jump;
      target = &at(resolved);
    }
    return resolved;
  }

  // RegexInstrBuilder.resolveJumps (:60-81)
  void resolve_jumps() {
    for (size_t i = 0; i < instrs_.size(); i++) {
      Instr in = instrs_[i];
      if (in.op == Op::Jump) {
        int r = resolved_jump(in.jump);
        if (r != -1) instrs_[i] = mk_jump(r, in.priority);
      } else if (in.op == Op::Split) {
        std::vector<int> res(in.split.size());
        for (size_t j = 0; j < in.split.size(); j++) {
          int r = resolved_jump(in.split[j]);
          res[j] = (r == -1) ? in.split[j] : r;
        }
        instrs_[i] = mk_split(res, in.priority);
      }
    }
  }

  // RegexInstrBuilder.createPartial (:107-209)
  void create_partial(const Node* ast) {
    if (!ast) throw std::logic_error("null node");
    switch (ast->kind) {
      case NodeKind::Concat:
        create_partial(ast->a);
        create_partial(ast->b);
        break;
      case NodeKind::Repetition: {
        int split_index = static_cast<int>(instrs_.size());
        instrs_.push_back(mk_null());
        create_partial(ast->a);
        instrs_.push_back(mk_jump(split_index, max_priority_));
        if (!leftmost_longest_) max_priority_++;
        int post = static_cast<int>(instrs_.size());
        instrs_[split_index] = mk_split({split_index + 1, post}, max_priority_);
        break;
      }
      case NodeKind::Counted: {
        int rep = 0;
        for (; rep < ast->min; rep++) create_partial(ast->a);
        std::vector<int> switches;
        for (; rep < ast->max; rep++) {
          switches.push_back(static_cast<int>(instrs_.size()));
          instrs_.push_back(mk_null());
          create_partial(ast->a);
        }
        int final_loc = static_cast<int>(instrs_.size());
        for (int loc : switches) instrs_[loc] = mk_split({loc + 1, final_loc}, max_priority_);
        break;
      }
      case NodeKind::Union: {
        int split_index = static_cast<int>(instrs_.size());
        if (!ast->a) throw std::logic_error("null node");
        if (ast->a->kind != NodeKind::Union) instrs_.push_back(mk_null());
        int first_target = static_cast<int>(instrs_.size());
        int first_priority = max_priority_;
        create_partial(ast->a);
        if (!leftmost_longest_ && ast->with_priority) max_priority_++;
        int first_jump = static_cast<int>(instrs_.size());
        instrs_.push_back(mk_null());
        int second_target = static_cast<int>(instrs_.size());
        create_partial(ast->b);
        if (!leftmost_longest_ && ast->with_priority) max_priority_++;
        int final_target = static_cast<int>(instrs_.size());
        instrs_[first_jump] = mk_jump(final_target, first_priority);
        std::vector<int> targets;
        if (at(first_target).op == Op::Split) {
          for (int t : at(first_target).split) targets.push_back(t);
        } else {
          targets.push_back(first_target);
        }
        if (second_target < static_cast<int>(instrs_.size()) && at(second_target).op == Op::Split) {
          for (int t : at(second_target).split) targets.push_back(t);
        } else {
          targets.push_back(second_target);
        }
        instrs_[split_index] = mk_split(targets, first_priority);
        break;
      }
      case NodeKind::Range:
        instrs_.push_back(mk_range(ast->range.start, ast->range.end, max_priority_));
        break;
      case NodeKind::Literal:
        for (char16_t c : ast->lit) instrs_.push_back(mk_range(c, c, max_priority_));
        break;
      default:
        throw std::logic_error("Unhandled ast node type");
    }
  }
};

}  // namespace

std::vector<Instr> build_program(const Node* ast, bool leftmost_longest) {
  return ProgramBuilder(leftmost_longest).build(ast);
}

// ------------------------------------------------------------------------------------------------
// DFA graph
// ------------------------------------------------------------------------------------------------
void DfaState::add_transition(CharRange r, DfaState* target) {
  bool added = false;
  for (size_t i = 0; i < transitions.size(); i++) {
    auto& ex = transitions[i];
    if (ex.first == r) {
      return;
    } else if (static_cast<int>(ex.first.end) + 1 == static_cast<int>(r.start)) {
      if (ex.second == target) {
        ex = {CharRange{ex.first.start, r.end}, target};
        added = true;
        break;
      }
    }
  }
  if (!added) {
    transitions.push_back({r, target});
    std::stable_sort(transitions.begin(), transitions.end(),
                     [](const auto& x, const auto& y) { return x.first.start < y.first.start; });
  }
}

DfaState* DfaState::step(uint16_t c) const {
  for (const auto& t : transitions)
    if (c >= t.first.start && c <= t.first.end) return t.second;
  return nullptr;
}

int DfaState::char_total() const {
  int total = 0;
  for (const auto& t : transitions) total += t.first.start;
  return total;
}

int Dfa::max_char() const {
  int mx = 0;
  for (const DfaState* s : states)
    for (const auto& t : s->transitions) {
      mx = std::max(mx, static_cast<int>(t.first.start));
      mx = std::max(mx, static_cast<int>(t.first.end));
    }
  return mx;
}

DfaState* Dfa::after(const std::u16string& str) const {
  DfaState* s = root();
  for (char16_t c : str) {
    s = s->step(c);
    if (!s) return nullptr;
  }
  return s;
}

void Dfa::prune_dead_states() {
  // DFA.findLiveStates (DFA.java:768-792)
  std::vector<char> live(states.size(), 0);
  live[root()->number] = 1;
  for (const DfaState* s : states)
    if (s->accepting) live[s->number] = 1;
  bool changed = true;
  while (changed) {
    changed = false;
    for (const DfaState* s : states) {
      if (!live[s->number]) {
        for (const auto& t : s->transitions)
          if (live[t.second->number]) {
            changed = true;
            live[s->number] = 1;
          }
      }
    }
  }
  for (DfaState* s : states) {
    std::vector<std::pair<CharRange, DfaState*>> keep;
    for (const auto& t : s->transitions)
      if (live[t.second->number]) keep.push_back(t);
    s->transitions.swap(keep);
  }
  std::vector<DfaState*> ns;
  for (DfaState* s : states)
    if (live[s->number]) ns.push_back(s);
  states.swap(ns);
  for (size_t i = 0; i < states.size(); i++) states[i]->number = static_cast<int>(i);
}

// ------------------------------------------------------------------------------------------------
// Subset construction
// ------------------------------------------------------------------------------------------------
namespace {

// StateSet.java
struct StateSet {
  struct Data {
    int distance, priority;
  };
  std::unordered_map<int, Data> starts;
  JIntSet states;
  bool seen_accepting = false;

  bool add(int s, int distance, int priority) {
    auto it = starts.find(s);
    if (it != starts.end()) {
      if (it->second.distance < distance) it->second = Data{distance, priority};
    } else {
      starts.emplace(s, Data{distance, priority});
    }
    return states.add(s);
  }
  int distance(int s) const { return starts.at(s).distance; }
  int priority(int s) const { return starts.at(s).priority; }
  // StateSet.prune (:37-53)
  bool prune(int accepting_state, int boundary, int priority) {
    bool removed = false;
    for (int s : states.items()) {
      if (s == accepting_state) continue;
      const Data& d = starts.at(s);
      if (d.distance < boundary || priority < d.priority) {
        states.remove(s);
        starts.erase(s);
        removed = true;
      }
    }
    return removed;
  }
  std::vector<int> key() const {
    std::vector<int> k = states.items();
    std::sort(k.begin(), k.end());
    return k;
  }
};

// CharRange.minimalCovering (CharRange.java:112-157)
std::vector<CharRange> minimal_covering(std::vector<CharRange> ranges) {
  if (ranges.size() < 2) return ranges;
  std::stable_sort(ranges.begin(), ranges.end(), [](const CharRange& a, const CharRange& b) { return a.start < b.start; });
  std::vector<CharRange> out;
  int last_start = -1, last_end = -1;
  for (size_t i = 0; i < ranges.size(); i++) {
    const CharRange cur = ranges[i];
    while (last_end < cur.end) {
      uint16_t start = cur.start, end = cur.end;
      if (last_start >= start) start = static_cast<uint16_t>(last_start + 1);
      if (last_end >= start) start = static_cast<uint16_t>(last_end + 1);
      for (size_t j = i + 1; j < ranges.size(); j++) {
        const CharRange& nx = ranges[j];
        if (nx.start > start && nx.start <= end) end = static_cast<uint16_t>(nx.start - 1);
        if (nx.end >= start && nx.end <= end) end = nx.end;
      }
      last_start = start;
      last_end = end;
      if (start > end) throw std::invalid_argument("Tried to create a character range with start larger than end");
      out.push_back(CharRange{start, end});
    }
  }
  std::stable_sort(out.begin(), out.end(), [](const CharRange& a, const CharRange& b) { return a.start < b.start; });
  return out;
}

// CharRange.coverAllChars (CharRange.java:81-110)
std::vector<CharRange> cover_all_chars(const std::vector<CharRange>& ranges) {
  std::vector<CharRange> all;
  if (ranges.empty()) {
    all.push_back(CharRange{0, 0xFFFF});
    return all;
  }
  bool have = false;
  CharRange cur{};
  for (const CharRange& r : ranges) {
    if (!have) {
      cur = r;
      have = true;
      if (cur.start > 0) all.push_back(CharRange{0, static_cast<uint16_t>(cur.start - 1)});
      all.push_back(cur);
    } else {
      if (r.start > static_cast<uint16_t>(cur.end + 1)) {
        uint16_t s = static_cast<uint16_t>(cur.end + 1), e = static_cast<uint16_t>(r.start - 1);
        if (s > e) throw std::invalid_argument("Tried to create a character range with start larger than end");
        all.push_back(CharRange{s, e});
      }
      all.push_back(r);
      cur = r;
    }
  }
  if (cur.end < 0xFFFF) all.push_back(CharRange{static_cast<uint16_t>(cur.end + 1), 0xFFFF});
  return all;
}

class SubsetBuilder {
 public:
  explicit SubsetBuilder(const std::vector<Instr>& prog) : prog_(prog), closure_cache_(prog.size()), cached_(prog.size(), 0) {}

  // NFAToDFACompiler._compile (:34-43)
  std::unique_ptr<Dfa> run(ConversionMode mode) {
    dfa_.reset(new Dfa());
    StateSet states;
    states.add(0, 0, 1);
    states = epsilon_closure(states);
    DfaState* root = dfa_->new_state(has_accepting(states), 0);
    states.seen_accepting = root->accepting;
    store(states, root);
    add_states(states, mode);
    return std::move(dfa_);
  }

 private:
  const std::vector<Instr>& prog_;
  std::unique_ptr<Dfa> dfa_;
  int next_state_ = 1;
  std::map<std::vector<int>, std::vector<std::pair<bool, DfaState*>>> sets_;  // key -> [(seenAccepting, dfa)]
  std::vector<std::vector<int>> closure_cache_;
  std::vector<char> cached_;

  bool is_accepting(int s) const { return prog_[s].op == Op::Match; }
  bool has_accepting(const StateSet& ss) const {
    for (int s : ss.states.items())
      if (is_accepting(s)) return true;
    return false;
  }

  // NFA.epsilonClosure (NFA.java:239-266); result in java HashSet iteration order
  const std::vector<int>& nfa_closure(int initial) {
    if (cached_[initial]) return closure_cache_[initial];
    JIntSet seen, closure;
    std::deque<int> pending;
    pending.push_back(initial);
    while (!pending.empty()) {
      int next = pending.front();
      pending.pop_front();
      seen.add(next);
      const Instr& in = prog_[next];
      if (in.op == Op::Split) {
        for (int t : in.split)
          if (!seen.contains(t)) pending.push_back(t);
      } else if (in.op == Op::Jump) {
        if (!seen.contains(in.jump)) pending.push_back(in.jump);
      } else {
        closure.add(next);
      }
    }
    closure_cache_[initial] = closure.items();
    cached_[initial] = 1;
    return closure_cache_[initial];
  }

  // NFAToDFACompiler.getEpsilonClosure (:138-155)
  StateSet epsilon_closure(const StateSet& states) {
    StateSet closure;
    for (int state : states.states.items()) {
      int priority = prog_[state].priority;
      for (int e : nfa_closure(state)) {
        if (is_accepting(e)) closure.seen_accepting = true;
        if (states.states.contains(e))
          closure.add(e, states.distance(e), priority);
        else
          closure.add(e, states.distance(state), priority);
      }
    }
    closure.seen_accepting |= states.seen_accepting;
    return closure;
  }

  // NFAToDFACompiler.transition (:168-178)
  StateSet transition(const StateSet& from, uint16_t c) {
    StateSet out;
    for (int state : from.states.items()) {
      const Instr& in = prog_[state];
      if (in.op == Op::CharRange && in.start <= c && in.end >= c) out.add(state + 1, from.distance(state) + 1, in.priority);
    }
    return out;
  }

  void store(const StateSet& ss, DfaState* d) { sets_[ss.key()].push_back({ss.seen_accepting, d}); }

  // NFAToDFACompiler.getDFA (:124-135)
  DfaState* lookup(const StateSet& ss) {
    auto it = sets_.find(ss.key());
    if (it == sets_.end()) return nullptr;
    for (auto& p : it->second)
      if (ss.states.size() == 1 || ss.seen_accepting == p.first) return p.second;
    return nullptr;
  }

  // NFAToDFACompiler.addNFAStatesToDFA (:56-122)
  void add_states(StateSet initial, ConversionMode mode) {
    std::vector<StateSet> pending;
    pending.push_back(std::move(initial));
    const bool searching = mode == ConversionMode::ContainedIn || mode == ConversionMode::DfaSearch;
    while (!pending.empty()) {
      StateSet states = std::move(pending.back());
      pending.pop_back();
      DfaState* dfa = lookup(states);
      if (!dfa) throw std::logic_error("state set without DFA state");
      StateSet closure = epsilon_closure(states);
      bool accepting = closure.seen_accepting;
      if (accepting && mode == ConversionMode::ContainedIn) continue;
      if (mode == ConversionMode::ContainedIn || (!accepting && mode == ConversionMode::DfaSearch)) closure.add(0, 0, 1);

      std::vector<CharRange> found;
      for (int s : closure.states.items())
        if (prog_[s].op == Op::CharRange) found.push_back(CharRange{prog_[s].start, prog_[s].end});
      std::vector<CharRange> ranges = cover_all_chars(minimal_covering(found));

      for (const CharRange& range : ranges) {
        StateSet post = epsilon_closure(transition(closure, range.start));
        if (!post.seen_accepting) post.seen_accepting = closure.seen_accepting || has_accepting(post);
        if (post.seen_accepting && searching) {
          bool removed;
          do {
            removed = false;
            for (int s : post.states.items()) {
              if (is_accepting(s)) {
                if (post.prune(s, post.distance(s), post.priority(s))) {
                  removed = true;
                  break;
                }
              }
            }
          } while (removed);
        }
        if (!post.seen_accepting && searching) {
          post.add(0, 0, 1);
          post = epsilon_closure(post);
        }
        DfaState* target = lookup(post);
        if (!target) {
          target = dfa_->new_state(has_accepting(post), next_state_++);
          store(post, target);
          pending.push_back(post);
        }
        dfa->add_transition(range, target);
      }
    }
  }
};

}  // namespace

std::unique_ptr<Dfa> subset_construction(const std::vector<Instr>& prog, ConversionMode mode) {
  return SubsetBuilder(prog).run(mode);
}

// ------------------------------------------------------------------------------------------------
// Minimisation
// ------------------------------------------------------------------------------------------------
// MinimizeDFA.java refines an initial partition (accepting?, #transitions, sum of range starts :113-150)
// with pairwise `equivalent` checks (:209-226) until a whole pass finds nothing to split (:60-95).
// That fixpoint is the coarsest partition stable under `equivalent`, which does not depend on the
// order groups are visited in, so a plain refinement loop reaches the same classes.  The numbering of
// the minimised states is fixed by the first-touch order of minimizeDFA (:23-29), reproduced below.
std::unique_ptr<Dfa> minimize(const Dfa& dfa) {
  const int n = dfa.count();
  std::vector<int> group(n, 0);
  {
    std::map<std::tuple<bool, size_t, int>, int> ids;
    for (const DfaState* s : dfa.states) {
      auto key = std::make_tuple(s->accepting, s->transitions.size(), s->char_total());
      auto it = ids.find(key);
      if (it == ids.end()) it = ids.emplace(key, static_cast<int>(ids.size())).first;
      group[s->number] = it->second;
    }
  }
  bool changed = true;
  while (changed) {
    changed = false;
    // signature of a state under the current partition
    std::map<std::pair<int, std::vector<int>>, int> ids;
    std::vector<int> next(n, 0);
    for (const DfaState* s : dfa.states) {
      std::vector<int> sig;
      sig.reserve(s->transitions.size() * 3);
      for (const auto& t : s->transitions) {
        sig.push_back(t.first.start);
        sig.push_back(t.first.end);
        sig.push_back(group[t.second->number]);
      }
      auto key = std::make_pair(group[s->number], std::move(sig));
      auto it = ids.find(key);
      if (it == ids.end()) it = ids.emplace(std::move(key), static_cast<int>(ids.size())).first;
      next[s->number] = it->second;
    }
    int before = *std::max_element(group.begin(), group.end()) + 1;
    if (static_cast<int>(ids.size()) != before) changed = true;
    group.swap(next);
  }

  std::unique_ptr<Dfa> out(new Dfa());
  std::unordered_map<int, DfaState*> by_group;
  int counter = 1;
  auto get = [&](const DfaState* orig) -> DfaState* {
    auto it = by_group.find(group[orig->number]);
    if (it != by_group.end()) return it->second;
    DfaState* m = out->states.empty() ? out->new_state(orig->accepting, 0) : out->new_state(orig->accepting, counter++);
    by_group[group[orig->number]] = m;
    return m;
  };
  for (const DfaState* orig : dfa.states) {
    DfaState* m = get(orig);
    for (const auto& t : orig->transitions) {
      DfaState* nx = get(t.second);
      m->add_transition(t.first, nx);
    }
  }
  return out;
}

std::unique_ptr<Dfa> compile_dfa(const std::vector<Instr>& prog, ConversionMode mode) {
  std::unique_ptr<Dfa> d = subset_construction(prog, mode);
  d->prune_dead_states();
  return minimize(*d);
}

// ------------------------------------------------------------------------------------------------
// Byte classes
// ------------------------------------------------------------------------------------------------
namespace {

// DFA.getNextEnd (DFA.java:529-546)
uint16_t next_end(const std::vector<CharRange>& all, size_t i, uint16_t next_start, const CharRange& left) {
  uint16_t ne = left.end;
  for (size_t j = i + 1; j < all.size(); j++) {
    const CharRange& right = all[j];
    if (right.end < next_start) continue;
    if (right.start > ne) return ne;
    if (next_start >= right.start)
      ne = std::min(ne, right.end);
    else
      ne = static_cast<uint16_t>(right.start - 1);
  }
  return ne;
}

// DFA.getDistinctCharRanges (DFA.java:509-527)
std::vector<CharRange> distinct_ranges(const std::vector<CharRange>& all) {
  std::vector<CharRange> out;
  uint16_t next_start = 0;
  for (size_t i = 0; i < all.size(); i++) {
    const CharRange& left = all[i];
    next_start = std::max(next_start, left.start);
    while (next_start <= left.end) {
      uint16_t ne = next_end(all, i, next_start, left);
      if (next_start > ne) throw std::invalid_argument("Tried to create a character range with start larger than end");
      out.push_back(CharRange{next_start, ne});
      if (ne == 0xFFFF) return out;
      next_start = static_cast<uint16_t>(ne + 1);
    }
  }
  return out;
}

}  // namespace

ByteClasses byte_classes(const Dfa& dfa) {
  // DFA.getSortedTransitions (:561-566)
  std::set<CharRange> uniq;
  for (const DfaState* s : dfa.states)
    for (const auto& t : s->transitions) uniq.insert(t.first);
  std::vector<CharRange> all(uniq.begin(), uniq.end());
  std::vector<CharRange> distinct = distinct_ranges(all);

  // DFA.charRanges + generateRangeGroups (:465-500): group distinct ranges by the set of
  // (state, target) pairs of the transitions they overlap.
  std::map<std::vector<std::pair<int, int>>, std::vector<CharRange>> groups;
  for (const CharRange& r : distinct) {
    std::vector<std::pair<int, int>> key;
    for (const DfaState* s : dfa.states)
      for (const auto& t : s->transitions)
        if (t.first.overlaps(r)) key.push_back({s->number, t.second->number});
    std::sort(key.begin(), key.end());
    key.erase(std::unique(key.begin(), key.end()), key.end());
    groups[key].push_back(r);
  }
  std::vector<std::vector<CharRange>> range_groups;
  for (auto& g : groups) {
    std::sort(g.second.begin(), g.second.end());
    range_groups.push_back(g.second);
  }
  // RangeGroup.compareTo (RangeGroup.java:10-27) is lexicographic order on the sorted range lists
  std::sort(range_groups.begin(), range_groups.end());

  // DFA.byteClasses (:438-463)
  ByteClasses bc;
  bc.ranges.assign(65537, 0);
  bc.n_groups = static_cast<int>(range_groups.size());
  bc.wide.assign(65536, 0);
  for (size_t g = 0; g < range_groups.size(); g++)
    for (const CharRange& r : range_groups[g])
      for (int i = r.start; i <= r.end; i++) bc.wide[i] = static_cast<uint16_t>(g + 1);
  int byte_class = 1;
  int catch_all = -1;  // ByteClasses.CATCHALL_INVALID
  for (const auto& g : range_groups) {
    if (g.back().end == 0xFFFF) catch_all = static_cast<int8_t>(byte_class);
    for (const CharRange& r : g) {
      int to = std::min(static_cast<int>(r.end) + 1, 65535);  // Arrays.fill(.., min(end + 1, 65535)): index 65535 is never written
      for (int i = r.start; i < to; i++) bc.ranges[i] = static_cast<uint8_t>(byte_class);
    }
    byte_class++;
    if (byte_class > 255) return bc;  // present == false
  }
  if (catch_all == -1) catch_all = 0;
  bc.present = true;
  bc.catch_all = catch_all;
  bc.byte_class_count = byte_class & 0xFF;
  return bc;
}

}  // namespace ndl
