// Host-side regex parser: regex string + flags -> AST.
//
// Restates needle-compiler/src/main/java/com/justinblank/strings/RegexParser.java (supported syntax,
// error surface, case folding, Unicode classes) and the factory logic of RegexAST/Union.java.  The
// reference is a stack machine over `Stack<Node>`; the same stack discipline is kept here because the
// shape of the resulting tree (which literals get merged, which unions are ordered) decides NFA
// instruction order and priorities further down.  java.util.HashSet iteration order is reproduced with
// JHashSet (jhash.h) where the reference builds unions by iterating a set.
#include <algorithm>
#include <climits>
#include <cstring>

#include "ast.h"
#include "jhash.h"
#include "needle_b200.h"

namespace ndl {

#include "unicode_tables.inc"

namespace {

std::vector<uint16_t> table_chars(const uint16_t (*t)[2], size_t n) {
  std::vector<uint16_t> out;
  for (size_t i = 0; i < n; i++)
    for (uint32_t c = t[i][0]; c <= t[i][1]; c++) out.push_back(static_cast<uint16_t>(c));
  return out;
}
#define TABLE_CHARS(T) table_chars(T, sizeof(T) / sizeof(T[0]))

struct CaseTables {
  uint16_t upper[65536], lower[65536], fold[65536];
  std::vector<std::vector<uint16_t>> by_fold;  // candidates (< 0xFFFF) grouped by fold value, ascending
  CaseTables() {
    for (uint32_t c = 0; c < 65536; c++) upper[c] = lower[c] = static_cast<uint16_t>(c);
    for (auto& p : kSimpleUpper) upper[p[0]] = p[1];
    for (auto& p : kSimpleLower) lower[p[0]] = p[1];
    by_fold.resize(65536);
    for (uint32_t c = 0; c < 65536; c++) {
      fold[c] = lower[upper[c]];  // Character.toLowerCase(Character.toUpperCase(c))
      if (c < 0xFFFF) by_fold[fold[c]].push_back(static_cast<uint16_t>(c));
    }
  }
};
const CaseTables& case_tables() {
  static const CaseTables t;
  return t;
}

// RegexParser.addCaseInsensitiveMatches (RegexParser.java:277-291)
void add_case_insensitive(JIntSet& chars, uint16_t c) {
  const CaseTables& t = case_tables();
  chars.add(c);
  uint16_t lower = t.fold[c];
  if (lower != t.upper[c]) {
    for (uint16_t cand : t.by_fold[lower]) chars.add(cand);
  }
}

struct RangeHash {  // Objects.hash(start, end) over boxed Characters (CharRange.java:231-233)
  uint32_t operator()(const CharRange& r) const { return 31u * (31u + r.start) + r.end; }
};
using RangeSet = JHashSet<CharRange, RangeHash>;

const uint16_t kHorizWs[] = {' ', '\t', 0x00A0, 0x1680, 0x180e, 0x2000, 0x2001, 0x2002, 0x2003, 0x2004,
                             0x2005, 0x2006, 0x2007, 0x2008, 0x2009, 0x200a, 0x202f, 0x205f, 0x3000};
const uint16_t kVertWs[] = {'\n', 0x000B, '\f', '\r', 0x0085, 0x2028, 0x2029};
const uint16_t kAsciiWs[] = {' ', '\t', '\n', 0x000B, '\f', '\r'};
template <size_t N>
std::vector<uint16_t> vec(const uint16_t (&a)[N]) {
  return std::vector<uint16_t>(a, a + N);
}

std::string narrow(const std::u16string& s) {
  std::string out;
  for (char16_t c : s) out.push_back(c < 128 ? static_cast<char>(c) : '?');
  return out;
}

class Parser {
 public:
  Parser(Ast& ast, const std::u16string& regex, int flags) : ast_(ast), regex_(regex) {
    dot_all_ = (flags & NDL_DOTALL) != 0;
    case_insensitive_ = (flags & NDL_CASE_INSENSITIVE) != 0;
    unicode_class_ = (flags & NDL_UNICODE_CHARACTER_CLASS) != 0;
    unicode_case_ = unicode_class_ ? true : (flags & NDL_UNICODE_CASE) != 0;  // RegexParser.java:72-78
  }

  Node* run();

 private:
  Ast& ast_;
  const std::u16string& regex_;
  size_t index_ = 0;
  int char_range_depth_ = 1;
  bool dot_all_, case_insensitive_, unicode_case_, unicode_class_;
  std::vector<Node*> nodes_;

  [[noreturn]] void error(const std::string& msg) const { throw SyntaxError(msg + ". Regex=" + narrow(regex_)); }
  uint16_t take() {
    if (index_ >= regex_.size()) throw std::out_of_range("charAt");  // StringIndexOutOfBoundsException
    return regex_[index_++];
  }
  bool peek_char(uint16_t c) const { return index_ < regex_.size() && regex_[index_] == c; }
  bool peek_string(const char* s) const {
    size_t n = std::strlen(s);
    if (index_ + n > regex_.size()) return false;
    for (size_t i = 0; i < n; i++)
      if (regex_[index_ + i] != static_cast<char16_t>(s[i])) return false;
    return true;
  }
  Node* pop() {
    if (nodes_.empty()) throw std::out_of_range("EmptyStackException");
    Node* n = nodes_.back();
    nodes_.pop_back();
    return n;
  }
  Node* peek() const {
    if (nodes_.empty()) throw std::out_of_range("EmptyStackException");
    return nodes_.back();
  }
  void assert_non_empty(const char* msg) const {
    if (nodes_.empty()) error(msg);
  }
  void reject_lazy_possessive() const {
    if (peek_char('?')) error("Reluctant quantifiers are not supported");
    if (peek_char('+')) error("Possessive quantifiers are not supported");
  }
  // RegexParser.concatenate (RegexParser.java:358-364)
  Node* concatenate(Node* next, Node* node) { return ast_.concatenate(next, node); }

  int consume_int();
  Node* parse_escape();
  Node* parse_octal();
  Node* parse_hex();
  Node* build_char_set();  // may return nullptr (Optional.empty)
  Node* with_alternate(Node* node, Node* alternate) {
    if (node) return alternate ? ast_.unordered(alternate, node) : node;
    return alternate;
  }
  Node* build_node(RangeSet& ranges, bool complemented);
  Node* build_ranges(RangeSet& ranges, bool complemented);
  void consume_named_group();
  void collapse_literals();
  void collapse_paren_nodes();
  Node* word_class();
};

// RegexParser._parse (RegexParser.java:100-275)
Node* Parser::run() {
  while (index_ < regex_.size()) {
    uint16_t c = take();
    switch (c) {
      case '.':
        if (dot_all_) {
          nodes_.push_back(ast_.range(0, 0xFFFF));
        } else {
          // everything except \n and \r (RegexParser.java:110)
          nodes_.push_back(ast_.unordered(
              ast_.range(0, 0x0009), ast_.unordered(ast_.range(0x000B, 0x000C), ast_.range(0x000E, 0xFFFF))));
        }
        break;
      case '^':
        error("'^' not supported yet");
      case '$':
        error("'$' not supported yet");
      case '(':
        nodes_.push_back(ast_.lparen());
        if (peek_string("?:")) {
          take();
          take();
        } else if (peek_string("?<")) {
          consume_named_group();
        }
        break;
      case '{': {
        if (nodes_.empty()) error("Found '{' with no preceding regex");
        int left = consume_int();
        uint16_t next = take();
        if (next == '}') {
          nodes_.push_back(ast_.counted(pop(), left, left));
          reject_lazy_possessive();
          break;
        } else if (next != ',') {
          error("Expected ',' in counted repetition");
        }
        int right = consume_int();
        nodes_.push_back(ast_.counted(pop(), left, right));
        next = take();
        if (next != '}') error("Found unclosed brackets");
        reject_lazy_possessive();
        break;
      }
      case '?':
        if (nodes_.empty()) error("");
        reject_lazy_possessive();
        nodes_.push_back(ast_.counted(pop(), 0, 1));
        break;
      case '[': {
        Node* set = build_char_set();
        if (set) nodes_.push_back(set);
        break;
      }
      case '+': {
        if (nodes_.empty()) error("Found '+' with no preceding regex");
        reject_lazy_possessive();
        Node* last = pop();
        nodes_.push_back(concatenate(last, ast_.repetition(last)));
        break;
      }
      case '*':
        if (nodes_.empty()) error("Found '*' with no preceding regex");
        reject_lazy_possessive();
        nodes_.push_back(ast_.repetition(pop()));
        break;
      case '|': {
        assert_non_empty("'|' cannot be the final character in a regex");
        collapse_literals();
        Node* last = pop();
        nodes_.push_back(ast_.ordered(last, nullptr));
        break;
      }
      case '\\':
        nodes_.push_back(parse_escape());
        break;
      case ')':
        collapse_paren_nodes();
        break;
      default:
        if (case_insensitive_) {
          if (unicode_case_) {
            JIntSet chars;
            add_case_insensitive(chars, c);
            if (chars.size() > 1) {
              Node* u = nullptr;
              for (int cc : chars.items()) {
                Node* lit = ast_.literal(static_cast<uint16_t>(cc));
                u = u ? ast_.unordered(u, lit) : lit;
              }
              nodes_.push_back(u);
            } else {
              nodes_.push_back(ast_.literal(c));
            }
          } else {
            Node* node = ast_.literal(c);
            if ('A' <= c && c <= 'Z') {
              nodes_.push_back(ast_.unordered(node, ast_.literal(static_cast<uint16_t>(c + 32))));
            } else if ('a' <= c && c <= 'z') {
              nodes_.push_back(ast_.unordered(node, ast_.literal(static_cast<uint16_t>(c - 32))));
            } else {
              nodes_.push_back(node);
            }
          }
        } else {
          nodes_.push_back(ast_.literal(c));
        }
    }
  }
  if (nodes_.empty()) return ast_.literal(std::u16string());
  Node* node = pop();
  if (node->kind == NodeKind::LParen) error("Unbalanced '(' found");
  while (!nodes_.empty()) {
    Node* next = pop();
    if (next->kind == NodeKind::Union && next->b == nullptr) {
      node = ast_.ordered(next->a, node);
    } else if (next->kind == NodeKind::Literal && node->kind == NodeKind::Literal) {
      node = ast_.literal(next->lit + node->lit);
    } else if (next->kind == NodeKind::LParen) {
      error("Unbalanced '(' found");
    } else {
      node = concatenate(next, node);
    }
  }
  return node;
}

// RegexParser.collapseLiterals (RegexParser.java:297-321)
void Parser::collapse_literals() {
  Node* last = pop();
  while (!nodes_.empty()) {
    Node* previous = peek();
    if (previous->kind != NodeKind::Union && previous->kind != NodeKind::LParen) {
      previous = pop();
      last = concatenate(previous, last);
    } else if (previous->kind == NodeKind::Union) {
      if (previous->b == nullptr) {
        pop();
        last = ast_.ordered(previous->a, last);
      } else {
        last = ast_.concatenate(pop(), last);
      }
    } else {
      break;  // LParen
    }
  }
  nodes_.push_back(last);
}

// RegexParser.collapseParenNodes (RegexParser.java:323-356)
void Parser::collapse_paren_nodes() {
  assert_non_empty("found unbalanced ')'");
  Node* node = nullptr;
  while (peek()->kind != NodeKind::LParen) {
    Node* previous = pop();
    if (node == nullptr) {
      node = previous;
    } else if (previous->kind == NodeKind::Union) {
      if (previous->a != nullptr && previous->b != nullptr) {
        node = ast_.concatenate(previous, node);
        continue;
      }
      assert_non_empty("found '|' with no preceding content");
      Node* next_next = peek();
      if (next_next->kind == NodeKind::LParen) {
        pop();
        nodes_.push_back(ast_.ordered(previous->a, node));
        return;
      }
      node = ast_.ordered(previous->a, node);
    } else {
      node = concatenate(previous, node);
    }
    assert_non_empty("found unbalanced ')'");
  }
  pop();
  if (node == nullptr) node = ast_.literal(std::u16string());  // "()|abc"
  nodes_.push_back(node);
}

// RegexParser.consumeInt (RegexParser.java:535-551)
int Parser::consume_int() {
  size_t initial = index_;
  while (index_ < regex_.size()) {
    uint16_t next = regex_[index_];
    if (next < '0' || next > '9') {
      if (index_ == initial) error("Expected number");
      long long v = 0;
      for (size_t i = initial; i < index_; i++) {
        v = v * 10 + (regex_[i] - '0');
        if (v > INT_MAX) error("Expected number");  // NumberFormatException
      }
      return static_cast<int>(v);
    }
    take();
  }
  error("Expected number");
}

Node* Parser::word_class() {
  return ast_.unordered(ast_.range('0', '9'),
                        ast_.unordered(ast_.range('_', '_'), ast_.unordered(ast_.range('a', 'z'), ast_.range('A', 'Z'))));
}

// RegexParser.parseEscapeSequence (RegexParser.java:372-529)
Node* Parser::parse_escape() {
  if (index_ >= regex_.size()) error("'\\' character with nothing following it");
  uint16_t c = take();
  switch (c) {
    case 'a': return ast_.range(0x0007, 0x0007);
    case 'A': error("\\A not supported yet");
    case 'B': error("\\B not supported yet");
    case 'b': error("\\b not supported yet");
    case 'c': error("\\c not supported yet");
    case 'd':
      return unicode_class_ ? ast_.of_chars(TABLE_CHARS(kUniDigit)) : ast_.range('0', '9');
    case 'D':
      return unicode_class_ ? ast_.complement_chars(TABLE_CHARS(kUniDigit))
                            : ast_.complement({CharRange{'0', '9'}});
    case 'e': return ast_.range(0x001B, 0x001B);
    case 'f': return ast_.range(0x000C, 0x000C);
    case 'G': error("\\G not supported yet");
    case 'H': return ast_.complement_chars(vec(kHorizWs));
    case 'h': return ast_.of_chars(vec(kHorizWs));
    case 'n': return ast_.range('\n', '\n');
    case 'p': error("\\p not supported yet");
    case 'r': return ast_.range('\r', '\r');
    case 's':
      return unicode_class_ ? ast_.of_chars(TABLE_CHARS(kUniSpace)) : ast_.of_chars(vec(kAsciiWs));
    case 'S':
      return unicode_class_ ? ast_.complement_chars(TABLE_CHARS(kUniSpace)) : ast_.complement_chars(vec(kAsciiWs));
    case 't': return ast_.range('\t', '\t');
    case 'w':
      return unicode_class_ ? ast_.of_chars(TABLE_CHARS(kUniWord)) : word_class();
    case 'W':
      if (unicode_class_) return ast_.complement_chars(TABLE_CHARS(kUniWord));
      return ast_.complement({CharRange{'0', '9'}, CharRange{'_', '_'}, CharRange{'a', 'z'}, CharRange{'A', 'Z'}});
    case 'x': return parse_hex();
    case 'V': return ast_.complement_chars(vec(kVertWs));
    case 'v': return ast_.of_chars(vec(kVertWs));
    case 'Z': error("\\Z not supported yet");
    case 'z': error("\\z not supported yet");
    case '0': return parse_octal();
    case '\\': return ast_.range('\\', '\\');
    case '[':
      // falls through to the metacharacter list when charRangeDepth <= 0, which never happens
      // (depth starts at 1) - same result either way
    case '|': case '(': case ')': case '$': case '*': case '?': case '+': case '{': case ':': case '^': case '.':
      return ast_.range(c, c);
    default: break;
  }
  if (c >= '1' && c <= '9') error("Backreferences are not supported");
  if (c < 'A') return ast_.range(c, c);
  if (c > 'Z' && c < 'a') return ast_.range(c, c);
  if (c > 'z') return ast_.range(c, c);
  error("Escape with unrecognized escaped character");
}

// RegexParser.parseOctal (RegexParser.java:553-569)
Node* Parser::parse_octal() {
  int count = 0;
  std::string str;
  auto peek_octal = [&]() { return index_ < regex_.size() && regex_[index_] >= '0' && regex_[index_] <= '7'; };
  while (count < 3 && peek_octal()) {
    if (count == 2 && str[0] > '3') break;
    str.push_back(static_cast<char>(take()));
    count++;
  }
  if (count == 0) error("Illegal octal escape");
  int v = 0;
  for (char ch : str) v = v * 8 + (ch - '0');
  return ast_.range(static_cast<uint16_t>(v), static_cast<uint16_t>(v));
}

// RegexParser.parseHexadecimal (RegexParser.java:571-584); only upper-case hex digits are accepted (:804-812)
Node* Parser::parse_hex() {
  int count = 0, v = 0;
  auto is_hex = [](uint16_t c) { return (c >= '0' && c <= '9') || (c >= 'A' && c <= 'F'); };
  while (count < 2 && index_ < regex_.size() && is_hex(regex_[index_])) {
    uint16_t c = take();
    v = v * 16 + (c <= '9' ? c - '0' : c - 'A' + 10);
    count++;
  }
  if (count != 2) error("Wrong number of hex chars");
  return ast_.range(static_cast<uint16_t>(v), static_cast<uint16_t>(v));
}

// RegexParser.buildCharSet (RegexParser.java:586-712)
Node* Parser::build_char_set() {
  char_range_depth_++;
  RangeSet ranges;
  bool has_last = false;
  uint16_t last = 0;
  size_t starting_index = index_;
  Node* alternate = nullptr;
  bool complemented = false;
  auto add_single = [&](uint16_t ch) { ranges.add(CharRange{ch, ch}); };
  while (index_ < regex_.size()) {
    uint16_t c = take();
    if (c == '^' && index_ == starting_index + 1) {
      complemented = true;
    } else if (c == ']') {
      if (has_last) add_single(last);
      char_range_depth_--;
      return with_alternate(build_node(ranges, complemented), alternate);
    } else if (c == '-') {
      if (index_ == regex_.size()) {
        error("Unterminated character range");
      } else if (peek_char(']')) {
        if (has_last) add_single(last);
        last = '-';
        has_last = true;
        continue;
      }
      if (!has_last) {
        last = c;
        has_last = true;
        continue;
      }
      uint16_t next = take();
      if (next == '\\') {
        if (peek_char('[') || peek_char(']') || peek_char('\\')) next = take();
      }
      if (next < last) error("Start of range must be less than the end");
      CharRange range{last, next};
      if (case_insensitive_) {
        if (unicode_case_) {
          JIntSet chars;
          for (uint32_t rc = last; rc <= next; rc++) {
            add_case_insensitive(chars, static_cast<uint16_t>(rc));
            // `char rangeChar <= next` never terminates in the reference when next == 0xFFFF; stop instead
            if (rc == 0xFFFF) break;
          }
          for (int rc : chars.items()) add_single(static_cast<uint16_t>(rc));
        } else {
          if (next < 'A' || 'z' < last) {
            ranges.add(range);
          } else {
            CharRange up{'A', 'Z'}, lo{'a', 'z'};
            if (range.overlaps(up)) {
              CharRange r{std::max(range.start, up.start), std::min(range.end, up.end)};
              ranges.add(r);
              ranges.add(CharRange{static_cast<uint16_t>(r.start + 32), static_cast<uint16_t>(r.end + 32)});
            }
            if (range.overlaps(lo)) {
              CharRange r{std::max(range.start, lo.start), std::min(range.end, lo.end)};
              ranges.add(r);
              ranges.add(CharRange{static_cast<uint16_t>(r.start - 32), static_cast<uint16_t>(r.end - 32)});
            }
            ranges.add(range);
          }
        }
      } else {
        ranges.add(range);
      }
      has_last = false;
    } else if (c == '[') {
      if (!ranges.empty()) {
        Node* range_node = nullptr;
        for (const CharRange& r : ranges.items()) {
          Node* rn = ast_.range(r);
          range_node = range_node ? ast_.unordered(range_node, rn) : rn;
        }
        ranges.clear();
        alternate = with_alternate(range_node, alternate);
      }
      Node* maybe = build_char_set();
      if (!maybe) error("Unbalanced [ token");
      if (peek_char(']')) {
        take();
        char_range_depth_--;
        return with_alternate(maybe, alternate);
      } else {
        alternate = with_alternate(maybe, alternate);
      }
    } else if (c == '\\') {
      if (peek_char('[') || peek_char(']') || peek_char('\\')) {
        uint16_t next = take();
        ranges.add(CharRange{next, next});
        last = next;
        has_last = true;
      } else {
        alternate = parse_escape();
      }
    } else {
      if (has_last) add_single(last);
      last = c;
      has_last = true;
    }
  }
  error("Parsing failed, unmatched [");
}

// RegexParser.buildNode / buildRanges (RegexParser.java:725-760)
Node* Parser::build_node(RangeSet& ranges, bool complemented) {
  if (ranges.empty()) return nullptr;
  if (ranges.size() == 1) {
    CharRange r = ranges.items()[0];
    if (complemented) return ast_.complement({r});
    return ast_.range(r);
  }
  return build_ranges(ranges, complemented);
}

Node* Parser::build_ranges(RangeSet& ranges, bool complemented) {
  // CharRange.compact (CharRange.java:201-219): stable sort by start, merge touching neighbours
  std::vector<CharRange> rs = ranges.items();
  std::stable_sort(rs.begin(), rs.end(), [](const CharRange& x, const CharRange& y) { return x.start < y.start; });
  std::vector<CharRange> sorted;
  CharRange current = rs[0];
  for (size_t i = 1; i < rs.size(); i++) {
    if (static_cast<uint16_t>(current.end + 1) == rs[i].start) {
      if (current.start > rs[i].end) throw std::invalid_argument("bad range");
      current = CharRange{current.start, rs[i].end};
    } else {
      sorted.push_back(current);
      current = rs[i];
    }
  }
  sorted.push_back(current);
  if (sorted.size() == 1) {
    if (complemented) return ast_.complement({sorted[0]});
    return ast_.range(sorted[0]);
  }
  if (complemented) return ast_.complement(sorted);
  Node* node = ast_.unordered(ast_.range(sorted[0]), ast_.range(sorted[1]));
  for (size_t i = 2; i < sorted.size(); i++) node = ast_.unordered(node, ast_.range(sorted[i]));
  return node;
}

// RegexParser.consumeNamedGroup (RegexParser.java:762-776).  The reference spins forever on a group
// name containing anything but [A-Za-z0-9]; that is reported as a syntax error here.
void Parser::consume_named_group() {
  size_t group_index = index_ + 2;
  while (group_index < regex_.size()) {
    uint16_t c = regex_[group_index];
    if (('A' <= c && c <= 'Z') || ('a' <= c && c <= 'z') || ('0' <= c && c <= '9')) {
      group_index++;
    } else if (c == '>') {
      index_ = group_index + 1;
      return;
    } else {
      error("Illegal character in group name");
    }
  }
}

}  // namespace

Node* Ast::of_chars(std::vector<uint16_t> chars) {
  if (chars.empty()) throw std::invalid_argument("Cannot create a union of zero characters");
  if (chars.size() == 1) return range(chars[0], chars[0]);
  std::sort(chars.begin(), chars.end());
  Node* u = nullptr;
  size_t start = 0;
  for (size_t i = 0; i < chars.size(); i++) {
    bool flush;
    if (i + 1 < chars.size())
      flush = chars[i] + 1 != chars[i + 1];
    else
      flush = true;
    if (flush) {
      Node* r = range(chars[start], chars[i]);
      u = u ? unordered(u, r) : r;
      start = i + 1;
    }
  }
  return u;
}

Node* Ast::complement(std::vector<CharRange> ranges) {
  if (ranges.empty()) throw std::invalid_argument("Can't complement empty set of ranges");
  std::sort(ranges.begin(), ranges.end());
  std::vector<Node*> out;
  bool have_last = false;
  CharRange last{};
  for (const CharRange& cur : ranges) {
    if (have_last) {
      uint16_t low = static_cast<uint16_t>(last.end + 1);
      uint16_t high = static_cast<uint16_t>(cur.start - 1);
      if (low <= high) out.push_back(range(low, high));
    } else {
      uint16_t high = static_cast<uint16_t>(cur.start - 1);  // wraps to 0xFFFF when the set contains \0
      out.push_back(range(0, high));
    }
    last = cur;
    have_last = true;
  }
  out.push_back(range(static_cast<uint16_t>(last.end + 1), 0xFFFF));
  Node* u = unordered(out.at(0), out.at(1));
  for (size_t i = 2; i < out.size(); i++) u = unordered(u, out[i]);
  return u;
}

Node* Ast::complement_chars(std::vector<uint16_t> chars) {
  if (chars.size() < 2) throw std::invalid_argument("Silly short complement");
  std::sort(chars.begin(), chars.end());
  std::vector<CharRange> rs;
  for (uint16_t c : chars) rs.push_back(CharRange{c, c});
  return complement(rs);
}

int Ast::min_length(const Node* n) {
  if (!n) throw std::logic_error("null node");
  switch (n->kind) {
    case NodeKind::Literal: return static_cast<int>(n->lit.size());
    case NodeKind::Range: return 1;
    case NodeKind::Union: return std::min(min_length(n->a), min_length(n->b));
    case NodeKind::Concat: return min_length(n->a) + min_length(n->b);
    case NodeKind::Repetition: return 0;
    case NodeKind::Counted: return n->min * min_length(n->a);
    case NodeKind::LParen: throw std::logic_error("UnsupportedOperationException");
  }
  return 0;
}

int Ast::max_length(const Node* n) {
  if (!n) throw std::logic_error("null node");
  switch (n->kind) {
    case NodeKind::Literal: return static_cast<int>(n->lit.size());
    case NodeKind::Range: return 1;
    case NodeKind::Union: {
      int l = max_length(n->a), r = max_length(n->b);
      return (l == kNoMax || r == kNoMax) ? kNoMax : std::max(l, r);
    }
    case NodeKind::Concat: {
      int l = max_length(n->a), r = max_length(n->b);
      return (l == kNoMax || r == kNoMax) ? kNoMax : l + r;
    }
    case NodeKind::Repetition: return kNoMax;
    case NodeKind::Counted: {
      int m = max_length(n->a);
      return m == kNoMax ? kNoMax : m * n->max;
    }
    case NodeKind::LParen: throw std::logic_error("UnsupportedOperationException");
  }
  return kNoMax;
}

Node* Ast::reversed(const Node* n) {
  if (!n) throw std::logic_error("null node");
  switch (n->kind) {
    case NodeKind::Literal: {
      std::u16string r(n->lit.rbegin(), n->lit.rend());
      return literal(r);
    }
    case NodeKind::Range: return const_cast<Node*>(n);
    case NodeKind::Union: return unordered(reversed(n->a), reversed(n->b));  // drops ordering (Union.java:46-48)
    case NodeKind::Concat: return concat_raw(reversed(n->b), reversed(n->a));
    case NodeKind::Repetition: return repetition(reversed(n->a));
    case NodeKind::Counted: return counted(reversed(n->a), n->min, n->max);
    case NodeKind::LParen: throw std::logic_error("UnsupportedOperationException");
  }
  return nullptr;
}

Node* parse_regex(Ast& ast, const std::u16string& regex, int flags) {
  try {
    return Parser(ast, regex, flags).run();
  } catch (const SyntaxError&) {
    throw;
  } catch (const std::exception& e) {
    // "Any other exception is a bug, wrap and rethrow" (RegexParser.java:94-97)
    throw SyntaxError(std::string("Unknown error while parsing regex '") + narrow(regex) + "': " + e.what());
  }
}

}  // namespace ndl
