"""Random regexes and matching haystacks: a Python restatement of the reference's test generator
(needle-compiler/src/test/java/com/justinblank/strings/RegexGenerator.java:23-62 generate / makeCharRangeNode,
:91-191 generateString / generateMinimalMatch) and of its printer (RegexAST/NodePrinter.java).  Test infrastructure.

Nodes are tuples: ("lit", str) ("range", lo, hi) ("cat", a, b) ("alt", a, b) ("star", a) ("rep", a, min, max).
The reference draws from java.util.Random without a seed; here numpy's generator is used with explicit seeds, so a
failure reproduces."""
import numpy as np

# RegexParserTest.ESCAPED_AS_LITERAL_CHARS (RegexParserTest.java:14)
ESCAPED_AS_LITERAL_CHARS = set("*()[$^+:?{")


class RegexGenerator:
    def __init__(self, rng: np.random.Generator, max_max_size: int):
        self.rng = rng
        self.max_size = int(rng.integers(0, max_max_size))
        self.count = 0

    def _next(self, bound):
        return int(self.rng.integers(0, bound))

    def _safe(self, c):
        return "B" if c in ESCAPED_AS_LITERAL_CHARS else c

    def _range(self):
        # makeCharRangeNode (:64-77)
        c1 = self._safe(chr(32 + self._next(128 - 32)))
        c2 = self._safe(chr(self._next(128 - ord(c1)) + ord(c1)))
        if c1 > c2:
            c2 = chr(ord(c1) + 1) if ord(c1) < 128 else c1
        return ("range", c1, c2)

    def generate(self):
        # generate (:23-62)
        if self.count + 1 >= self.max_size:
            return self._range()
        self.count += 1
        kind = self._next(8)
        if kind == 0:
            child = self.generate()
            i1 = self._next(7)
            i2 = 0 if i1 == 0 else self._next(i1)
            return ("rep", child, i2, i1)
        if kind == 1:
            return ("star", self.generate())
        if kind == 2:
            a = self.generate()
            return ("alt", a, self.generate())
        if kind in (3, 4, 5):
            a = self.generate()
            return ("cat", a, self.generate())
        if kind == 6:
            return ("lit", self._safe(chr(self._next(128))))
        return self._range()

    def generate_string(self, node) -> str:
        # addToString (:134-180)
        k = node[0]
        if k == "lit":
            return node[1]
        if k == "range":
            lo, hi = ord(node[1]), ord(node[2])
            return node[1] if lo == hi else chr(lo + self._next(hi - lo))
        if k == "cat":
            return self.generate_string(node[1]) + self.generate_string(node[2])
        if k == "alt":
            return self.generate_string(node[1] if self._next(2) else node[2])
        if k == "star":
            return "".join(self.generate_string(node[1]) for _ in range(self._next(8)))
        _, child, mn, mx = node
        count = mx if mx == mn else mn + self._next(mx - mn)
        return "".join(self.generate_string(child) for _ in range(count))


def _escape(c):
    return "\\" + c if c in "*?+(){[$^:|\\" else c


def _needs_parens(parent, child):
    if child[0] == "range":
        return False
    if child[0] == "lit" and len(child[1]) == 1:
        return False
    if parent[0] == "cat":
        return child[0] == "alt"
    return True


def print_node(node) -> str:
    """NodePrinter.print: the regex source of a node (with the printer's redundant parentheses)."""
    k = node[0]
    if k == "range":
        lo, hi = node[1], node[2]
        if lo == hi:
            return _escape(lo)
        esc = lambda c: "\\" + c if c in "[]\\" else c  # noqa: E731
        return "[" + esc(lo) + "-" + esc(hi) + "]"
    if k == "lit":
        return "".join(_escape(c) for c in node[1])

    def child(c):
        s = print_node(c)
        return "(" + s + ")" if _needs_parens(node, c) else s
    if k == "cat":
        return child(node[1]) + child(node[2])
    if k == "alt":
        return child(node[1]) + "|" + child(node[2])
    if k == "star":
        return child(node[1]) + "*"
    _, c, mn, mx = node
    return child(c) + ("?" if (mn, mx) == (0, 1) else "{%d,%d}" % (mn, mx))


def has_choice(node) -> bool:
    """Does the regex contain a choice whose outcome java.util.regex / Python `re` decide by priority (an alternation, or a
    repetition that may stop early): the only place where leftmost-first and needle's automaton can pick different ends."""
    k = node[0]
    if k in ("lit", "range"):
        return False
    if k == "alt" or k == "star":
        return True
    if k == "rep":
        return node[2] != node[3] or has_choice(node[1])
    return has_choice(node[1]) or has_choice(node[2])
