// Regex AST for the host-side compiler (regex -> NFA -> 4 DFAs -> table blob).
//
// Restates the node kinds and the length algebra of the reference's
// needle-compiler/src/main/java/com/justinblank/strings/RegexAST/{Node,Union,Concatenation,LiteralNode,
// CharRangeNode,Repetition,CountedRepetition,LParenNode}.java.  Node identity matters in the reference
// (Union.of dedups with Object.equals, LiteralNode is mutated in place by append), so nodes live in an
// arena and are compared by pointer exactly where the reference compares by reference.
#pragma once
#include <cstdint>
#include <deque>
#include <stdexcept>
#include <string>
#include <vector>

namespace ndl {

// PatternSyntaxException (RegexParser.java:91-97, 293-295)
struct SyntaxError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// PatternClassCompilationException (DFACompiler.java:34-36, 71-73)
struct CompileError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// IllegalStateException for > 16383 states (DFACompiler.java:76-83)
struct TooLargeError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
// IllegalArgumentException for unknown flag bits (CompilerOptions.java:10-12)
struct FlagsError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct CharRange {
  uint16_t start = 0, end = 0;
  bool operator==(const CharRange& o) const { return start == o.start && end == o.end; }
  bool operator!=(const CharRange& o) const { return !(*this == o); }
  // CharRange.compareTo (CharRange.java:236-247)
  bool operator<(const CharRange& o) const { return start != o.start ? start < o.start : end < o.end; }
  bool overlaps(const CharRange& o) const { return start <= o.end && end >= o.start; }
  bool single() const { return start == end; }
};

enum class NodeKind { Literal, Range, Union, Concat, Repetition, Counted, LParen };

struct Node {
  NodeKind kind;
  std::u16string lit;          // Literal
  CharRange range;             // Range
  Node* a = nullptr;           // Union.left / Concat.head / Repetition.node / Counted.node
  Node* b = nullptr;           // Union.right (may be null while parsing) / Concat.tail
  bool with_priority = false;  // Union
  int min = 0, max = 0;        // Counted
};

constexpr int kNoMax = -1;  // Optional.empty() for maxLength

class Ast {
 public:
  Node* literal(const std::u16string& s) {
    Node* n = mk(NodeKind::Literal);
    n->lit = s;
    return n;
  }
  Node* literal(uint16_t c) { return literal(std::u16string(1, static_cast<char16_t>(c))); }
  Node* range(uint16_t s, uint16_t e) {
    // CharRange constructor (CharRange.java:19-26)
    if (s > e) throw std::invalid_argument("Tried to create a character range with start larger than end");
    Node* n = mk(NodeKind::Range);
    n->range = CharRange{s, e};
    return n;
  }
  Node* range(CharRange r) { return range(r.start, r.end); }
  Node* concat_raw(Node* h, Node* t) {
    if (!h || !t) throw std::invalid_argument("Cannot concatenate nothing");
    Node* n = mk(NodeKind::Concat);
    n->a = h;
    n->b = t;
    return n;
  }
  Node* repetition(Node* x) {
    if (!x) throw std::invalid_argument("Cannot repeat nothing");
    Node* n = mk(NodeKind::Repetition);
    n->a = x;
    return n;
  }
  Node* counted(Node* x, int mn, int mx) {
    if (!x) throw std::invalid_argument("Cannot repeat nothing");
    if (mn > mx) throw std::invalid_argument("Repetition range is invalid");
    Node* n = mk(NodeKind::Counted);
    n->a = x;
    n->min = mn;
    n->max = mx;
    return n;
  }
  Node* lparen() {
    if (!lparen_) lparen_ = mk(NodeKind::LParen);
    return lparen_;
  }
  Node* union_raw(Node* l, Node* r, bool prio) {
    if (!l) throw std::invalid_argument("Cannot union nothing");
    Node* n = mk(NodeKind::Union);
    n->a = l;
    n->b = r;
    n->with_priority = prio;
    return n;
  }

  // Object.equals / LiteralNode.equals (LiteralNode.java:63-68).  A null receiver is the reference's
  // NullPointerException, which RegexParser.parse wraps into a PatternSyntaxException.
  static bool equals(const Node* x, const Node* y) {
    if (!x) throw std::logic_error("null node");
    if (x->kind == NodeKind::Literal) return y && y->kind == NodeKind::Literal && x->lit == y->lit;
    return x == y;
  }

  // Union.of (Union.java:58-75)
  Node* union_of(Node* l, Node* r, bool prio) {
    if (equals(l, r)) return l;
    if (l->kind == NodeKind::Union) {
      if (equals(l->a, r) || equals(l->b, r)) return l;
    }
    if (r && r->kind == NodeKind::Union) {
      if (equals(r->a, l) || equals(r->b, l)) return r;
    }
    return union_raw(l, r, prio);
  }
  Node* ordered(Node* l, Node* r) { return union_of(l, r, true); }
  Node* unordered(Node* l, Node* r) { return union_of(l, r, false); }

  // Concatenation.concatenate (Concatenation.java:21-36); mutates a literal head in place.
  Node* concatenate(Node* h, Node* t) {
    if (h && t && h->kind == NodeKind::Literal && t->kind == NodeKind::Literal) {
      h->lit += t->lit;
      return h;
    }
    if (h && t && h->kind == NodeKind::Literal && t->kind == NodeKind::Concat && t->a->kind == NodeKind::Literal) {
      h->lit += t->a->lit;
      return concat_raw(h, t->b);
    }
    return concat_raw(h, t);
  }

  // Union.ofChars (Union.java:77-113): sorted chars -> union of maximal runs.
  Node* of_chars(std::vector<uint16_t> chars);
  // Union.complement(List<CharRangeNode>) (Union.java:115-150), char arithmetic wraps like Java's.
  Node* complement(std::vector<CharRange> ranges);
  // Union.complement(String) (Union.java:152-164)
  Node* complement_chars(std::vector<uint16_t> chars);

  // Node.minLength / maxLength / reversed for each kind.
  static int min_length(const Node* n);
  static int max_length(const Node* n);  // kNoMax when unbounded
  Node* reversed(const Node* n);

 private:
  Node* mk(NodeKind k) {
    nodes_.emplace_back();
    nodes_.back().kind = k;
    return &nodes_.back();
  }
  std::deque<Node> nodes_;
  Node* lparen_ = nullptr;
};

// RegexParser.parse(regex, flags) (RegexParser.java:86-98).  Throws SyntaxError.
Node* parse_regex(Ast& ast, const std::u16string& regex, int flags);

}  // namespace ndl
