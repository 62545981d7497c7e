// Iteration-order-faithful stand-in for java.util.HashSet.
//
// The reference's subset construction iterates java.util.HashSet<Integer> in several places where the
// visiting order decides a tie (StateSet.add keeps the *first* priority seen at equal distance:
// needle-compiler/.../StateSet.java:16-27, driven from NFAToDFACompiler.java:139-151 and
// NFA.java:239-266), and the parser iterates HashSet<Character>/HashSet<CharRange> when it builds
// unions (RegexParser.java:214-226, 637-644, 668-677).  To produce the same automata - and the same
// state numbering as the snapshot fixtures - this container reproduces OpenJDK's HashMap layout:
// power-of-two table, index = (h ^ h>>>16) & (cap-1), insertion order inside a bucket, doubling at
// size > 0.75*cap with order-preserving bucket splits, and the "chain reaches 9 while cap < 64 =>
// resize" rule of treeifyBin.  Tree bins (cap >= 64 and 9 colliding keys) are not modelled; they
// keep list order for the keys already present and are not reachable with the small dense integer
// keys used here.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace ndl {

template <class K, class Hash>
class JHashSet {
 public:
  JHashSet() = default;
  // java.util.HashSet(Collection) sizing: new HashMap<>(max((int)(n/.75f)+1, 16))
  static JHashSet withExpected(size_t n) {
    JHashSet s;
    int cap = static_cast<int>(static_cast<float>(n) / .75f) + 1;
    if (cap < 16) cap = 16;
    s.threshold_ = tableSizeFor(cap);
    return s;
  }

  bool add(const K& k) {
    if (tab_.empty()) resize();
    uint32_t h = spread(Hash()(k));
    std::vector<K>& b = tab_[h & (tab_.size() - 1)];
    for (const K& e : b)
      if (e == k) return false;
    b.push_back(k);
    if (b.size() >= 9 && tab_.size() < 64) resize();  // treeifyBin on a small table resizes instead
    if (++size_ > threshold_) resize();
    return true;
  }

  bool contains(const K& k) const {
    if (tab_.empty()) return false;
    uint32_t h = spread(Hash()(k));
    const std::vector<K>& b = tab_[h & (tab_.size() - 1)];
    for (const K& e : b)
      if (e == k) return true;
    return false;
  }

  bool remove(const K& k) {
    if (tab_.empty()) return false;
    uint32_t h = spread(Hash()(k));
    std::vector<K>& b = tab_[h & (tab_.size() - 1)];
    for (size_t i = 0; i < b.size(); i++)
      if (b[i] == k) {
        b.erase(b.begin() + i);
        size_--;
        return true;
      }
    return false;
  }

  // HashMap.clear() keeps the table length.
  void clear() {
    for (auto& b : tab_) b.clear();
    size_ = 0;
  }

  size_t size() const { return size_; }
  bool empty() const { return size_ == 0; }

  // Snapshot in java iteration order (bucket index ascending, insertion order within a bucket).
  std::vector<K> items() const {
    std::vector<K> out;
    out.reserve(size_);
    for (const auto& b : tab_)
      for (const K& e : b) out.push_back(e);
    return out;
  }

 private:
  static uint32_t spread(uint32_t h) { return h ^ (h >> 16); }
  static size_t tableSizeFor(int cap) {
    size_t n = 1;
    while (n < static_cast<size_t>(cap)) n <<= 1;
    return n;
  }
  void resize() {
    size_t oldCap = tab_.size();
    size_t newCap, newThr;
    if (oldCap > 0) {
      newCap = oldCap << 1;
      newThr = (oldCap >= 16) ? (threshold_ << 1) : static_cast<size_t>(newCap * 0.75f);
    } else if (threshold_ > 0) {
      newCap = threshold_;
      newThr = static_cast<size_t>(newCap * 0.75f);
    } else {
      newCap = 16;
      newThr = 12;
    }
    std::vector<std::vector<K>> nt(newCap);
    for (auto& b : tab_)
      for (const K& e : b) nt[spread(Hash()(e)) & (newCap - 1)].push_back(e);  // order-preserving split
    tab_.swap(nt);
    threshold_ = newThr;
  }

  std::vector<std::vector<K>> tab_;
  size_t size_ = 0;
  size_t threshold_ = 0;
};

struct JIntHash {
  uint32_t operator()(int v) const { return static_cast<uint32_t>(v); }  // Integer.hashCode / Character.hashCode
};
using JIntSet = JHashSet<int, JIntHash>;

}  // namespace ndl
