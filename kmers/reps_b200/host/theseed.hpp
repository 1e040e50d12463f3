// theseed.hpp -- C++17 host-side mirror of the reference's Java classes for the protein k-mer path.
//
// The reference's host language is Java; this image has no JDK, so the host side above the C ABI
// (include/kmersreps.h) is written in C++ with the reference's names, argument meaning and error
// behaviour, so that callers (tools.cpp) read like the reference's processors:
//
//   org.theseed.sequence.Sequence / FastaInputStream / FastaOutputStream
//   org.theseed.sequence.SequenceKmers / ProteinKmers      (similarity, distance, size, INFINITY, static K)
//   org.theseed.proteins.kmers.reps.RepGenome / RepGenomeDb / RepGenomeDb.Representation
//   org.theseed.proteins.kmers.KmerCollectionGroup
//
// All of them live in the un-vendored org.theseed:shared:1.0.0 (reference pom.xml:80-94); each member
// below cites the reference call site that fixes its meaning.  Similarities are ALWAYS computed by
// libkmersreps_cuda.so; there is no CPU path here.  Batched overloads (findClosest(vector), checkGenomes,
// SimilarityMatrix) are the additions the rewritten loops call.
#pragma once
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <regex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../../include/kmersreps.h"

namespace theseed {

// ---------------------------------------------------------------------------------------------------
// java.* behaviours the outputs depend on
namespace java {

inline uint32_t string_hash(const std::string &s) {   // String.hashCode()
    uint32_t h = 0;
    for (unsigned char c : s) h = 31u * h + c;
    return h;
}

// Iteration order of java.util.HashMap<String,V> (OpenJDK 8+): RepGenomeDb stores its representatives in
// one keyed by genome id, so save(), every `for (RepGenome rep : repDb)` loop and the tie-breaking of
// findClosest inherit this order (data/rep50.db is in exactly this order).
//
// This is a structural emulation of java.util.HashMap for String keys, including tree bins:
// a bin whose 9th node arrives at table capacity >= 64 becomes a red-black tree ordered by (signed spread hash,
// String.compareTo); its iteration order is the `next` chain, which treeify()/putTreeVal()/moveRootToFront()/
// split()/untreeify()/removeTreeNode() maintain exactly as OpenJDK's HashMap.java does (restated from the OpenJDK
// source, which is not part of the reference tree; String keys never need tieBreakOrder(), so the order is
// deterministic).  Nodes are indexed by insertion number; -1 is null.
class HashMapOrder {
  public:
    // put(key, v): returns the insertion index of the key (existing or new); *is_new tells which
    int put(const std::string &key, bool *is_new = nullptr) {
        auto it = index_.find(key);
        if (it != index_.end()) { if (is_new) *is_new = false; return it->second; }   // existing mapping keeps its place
        if (is_new) *is_new = true;
        const int x = (int)nd_.size();
        uint32_t h = string_hash(key); h ^= h >> 16;
        nd_.push_back(Node{}); nd_[x].hash = (int32_t)h; keys_.push_back(key);
        index_[key] = x;
        if (tab_.empty()) { tab_.assign(16, -1); threshold_ = 12; }
        const size_t i = h & (tab_.size() - 1);
        int p = tab_[i];
        if (p < 0) tab_[i] = x;
        else if (nd_[p].tree) putTreeVal(p, x);
        else {
            for (int binCount = 0;; ++binCount) {
                const int e = nd_[p].next;
                if (e < 0) {
                    nd_[p].next = x;
                    if (binCount >= TREEIFY_THRESHOLD - 1) treeifyBin(h);
                    break;
                }
                p = e;
            }
        }
        if (++size_ > threshold_) resize();
        return x;
    }
    bool contains(const std::string &key) const { return index_.count(key) != 0; }
    int find(const std::string &key) const { auto it = index_.find(key); return it == index_.end() ? -1 : it->second; }
    // Iterator.remove() (movable = false, UpdateProcessor.java:96-103) or Map.remove(key) (movable = true)
    void remove(const std::string &key, bool movable = false) {
        auto it = index_.find(key);
        if (it == index_.end()) return;
        const int node = it->second;
        const size_t i = (uint32_t)nd_[node].hash & (tab_.size() - 1);
        if (nd_[node].tree) removeTreeNode(node, movable);
        else if (tab_[i] == node) tab_[i] = nd_[node].next;
        else { int p = tab_[i]; while (nd_[p].next != node) p = nd_[p].next; nd_[p].next = nd_[node].next; }
        index_.erase(it); --size_;
    }
    size_t size() const { return size_; }
    std::vector<int> order() const {   // insertion indices in iteration order
        std::vector<int> o; o.reserve(size_);
        for (int head : tab_) for (int e = head; e >= 0; e = nd_[e].next) o.push_back(e);
        return o;
    }
    size_t treeBins() const { size_t n = 0; for (int head : tab_) n += head >= 0 && nd_[head].tree; return n; }
    // structural self-check used by the tests: red-black and ordering invariants of every tree bin, and that each
    // bin's next chain holds exactly the nodes of its tree
    bool checkInvariants() const {
        for (size_t i = 0; i < tab_.size(); ++i) {
            int head = tab_[i];
            if (head < 0 || !nd_[head].tree) { for (int e = head; e >= 0; e = nd_[e].next) if (((uint32_t)nd_[e].hash & (tab_.size() - 1)) != i) return false; continue; }
            int root = head; while (nd_[root].parent >= 0) root = nd_[root].parent;
            size_t chain = 0; int prev = -1;
            for (int e = head; e >= 0; prev = e, e = nd_[e].next) { if (nd_[e].prev != prev || !nd_[e].tree) return false; ++chain; }
            size_t cnt = 0; int bh = -1;
            if (nd_[root].red || !checkTree(root, 0, &bh, &cnt) || cnt != chain) return false;
        }
        return true;
    }
  private:
    static constexpr int TREEIFY_THRESHOLD = 8, UNTREEIFY_THRESHOLD = 6, MIN_TREEIFY_CAPACITY = 64;
    struct Node { int32_t hash = 0; int next = -1, prev = -1, parent = -1, left = -1, right = -1; bool red = false, tree = false; };

    // order inside a tree bin: signed hash, then String.compareTo (never 0 for two distinct keys)
    int dirOf(int x, int p) const {
        if (nd_[p].hash > nd_[x].hash) return -1;
        if (nd_[p].hash < nd_[x].hash) return 1;
        return keys_[x].compare(keys_[p]) < 0 ? -1 : 1;
    }
    int rootOf(int r) const { while (nd_[r].parent >= 0) r = nd_[r].parent; return r; }

    void treeifyBin(uint32_t hash) {
        if (tab_.size() < (size_t)MIN_TREEIFY_CAPACITY) { resize(); return; }
        const size_t index = hash & (tab_.size() - 1);
        int tl = -1;
        for (int e = tab_[index]; e >= 0; e = nd_[e].next) {   // replacementTreeNode keeps the list order
            Node &n = nd_[e]; n.tree = true; n.prev = tl; n.parent = n.left = n.right = -1; n.red = false; tl = e;
        }
        treeify(tab_[index]);
    }
    void treeify(int head) {
        int root = -1;
        for (int x = head, next; x >= 0; x = next) {
            next = nd_[x].next;
            nd_[x].left = nd_[x].right = -1;
            if (root < 0) { nd_[x].parent = -1; nd_[x].red = false; root = x; continue; }
            for (int p = root;;) {
                const int dir = dirOf(x, p), xp = p;
                p = dir <= 0 ? nd_[p].left : nd_[p].right;
                if (p < 0) {
                    nd_[x].parent = xp;
                    if (dir <= 0) nd_[xp].left = x; else nd_[xp].right = x;
                    root = balanceInsertion(root, x);
                    break;
                }
            }
        }
        moveRootToFront(root);
    }
    void untreeify(int head) { for (int q = head; q >= 0; q = nd_[q].next) { Node &n = nd_[q]; n.tree = false; n.prev = n.parent = n.left = n.right = -1; n.red = false; } }
    void putTreeVal(int first, int x) {
        const int root = nd_[first].parent >= 0 ? rootOf(first) : first;
        for (int p = root;;) {
            const int dir = dirOf(x, p), xp = p;
            p = dir <= 0 ? nd_[p].left : nd_[p].right;
            if (p < 0) {
                const int xpn = nd_[xp].next;
                Node &n = nd_[x]; n.tree = true; n.next = xpn;
                if (dir <= 0) nd_[xp].left = x; else nd_[xp].right = x;
                nd_[xp].next = x;
                n.parent = n.prev = xp;
                if (xpn >= 0) nd_[xpn].prev = x;
                moveRootToFront(balanceInsertion(root, x));
                return;
            }
        }
    }
    void moveRootToFront(int root) {
        if (root < 0 || tab_.empty()) return;
        const size_t index = (uint32_t)nd_[root].hash & (tab_.size() - 1);
        const int first = tab_[index];
        if (root == first) return;
        tab_[index] = root;
        const int rp = nd_[root].prev, rn = nd_[root].next;
        if (rn >= 0) nd_[rn].prev = rp;
        if (rp >= 0) nd_[rp].next = rn;
        if (first >= 0) nd_[first].prev = root;
        nd_[root].next = first;
        nd_[root].prev = -1;
    }
    int rotateLeft(int root, int p) {
        int r, pp, rl;
        if (p >= 0 && (r = nd_[p].right) >= 0) {
            if ((rl = nd_[p].right = nd_[r].left) >= 0) nd_[rl].parent = p;
            if ((pp = nd_[r].parent = nd_[p].parent) < 0) { root = r; nd_[r].red = false; }
            else if (nd_[pp].left == p) nd_[pp].left = r;
            else nd_[pp].right = r;
            nd_[r].left = p; nd_[p].parent = r;
        }
        return root;
    }
    int rotateRight(int root, int p) {
        int l, pp, lr;
        if (p >= 0 && (l = nd_[p].left) >= 0) {
            if ((lr = nd_[p].left = nd_[l].right) >= 0) nd_[lr].parent = p;
            if ((pp = nd_[l].parent = nd_[p].parent) < 0) { root = l; nd_[l].red = false; }
            else if (nd_[pp].right == p) nd_[pp].right = l;
            else nd_[pp].left = l;
            nd_[l].right = p; nd_[p].parent = l;
        }
        return root;
    }
    int balanceInsertion(int root, int x) {
        nd_[x].red = true;
        for (int xp, xpp, xppl, xppr;;) {
            if ((xp = nd_[x].parent) < 0) { nd_[x].red = false; return x; }
            if (!nd_[xp].red || (xpp = nd_[xp].parent) < 0) return root;
            if (xp == (xppl = nd_[xpp].left)) {
                if ((xppr = nd_[xpp].right) >= 0 && nd_[xppr].red) { nd_[xppr].red = false; nd_[xp].red = false; nd_[xpp].red = true; x = xpp; }
                else {
                    if (x == nd_[xp].right) { root = rotateLeft(root, x = xp); xpp = (xp = nd_[x].parent) < 0 ? -1 : nd_[xp].parent; }
                    if (xp >= 0) { nd_[xp].red = false; if (xpp >= 0) { nd_[xpp].red = true; root = rotateRight(root, xpp); } }
                }
            } else {
                if (xppl >= 0 && nd_[xppl].red) { nd_[xppl].red = false; nd_[xp].red = false; nd_[xpp].red = true; x = xpp; }
                else {
                    if (x == nd_[xp].left) { root = rotateRight(root, x = xp); xpp = (xp = nd_[x].parent) < 0 ? -1 : nd_[xp].parent; }
                    if (xp >= 0) { nd_[xp].red = false; if (xpp >= 0) { nd_[xpp].red = true; root = rotateLeft(root, xpp); } }
                }
            }
        }
    }
    bool isRed(int n) const { return n >= 0 && nd_[n].red; }
    int balanceDeletion(int root, int x) {
        for (int xp, xpl, xpr;;) {
            if (x < 0 || x == root) return root;
            if ((xp = nd_[x].parent) < 0) { nd_[x].red = false; return x; }
            if (nd_[x].red) { nd_[x].red = false; return root; }
            if ((xpl = nd_[xp].left) == x) {
                if ((xpr = nd_[xp].right) >= 0 && nd_[xpr].red) {
                    nd_[xpr].red = false; nd_[xp].red = true;
                    root = rotateLeft(root, xp);
                    xpr = (xp = nd_[x].parent) < 0 ? -1 : nd_[xp].right;
                }
                if (xpr < 0) x = xp;
                else {
                    int sl = nd_[xpr].left, sr = nd_[xpr].right;
                    if (!isRed(sr) && !isRed(sl)) { nd_[xpr].red = true; x = xp; }
                    else {
                        if (!isRed(sr)) {
                            if (sl >= 0) nd_[sl].red = false;
                            nd_[xpr].red = true;
                            root = rotateRight(root, xpr);
                            xpr = (xp = nd_[x].parent) < 0 ? -1 : nd_[xp].right;
                        }
                        if (xpr >= 0) { nd_[xpr].red = xp < 0 ? false : nd_[xp].red; if ((sr = nd_[xpr].right) >= 0) nd_[sr].red = false; }
                        if (xp >= 0) { nd_[xp].red = false; root = rotateLeft(root, xp); }
                        x = root;
                    }
                }
            } else {
                if (xpl >= 0 && nd_[xpl].red) {
                    nd_[xpl].red = false; nd_[xp].red = true;
                    root = rotateRight(root, xp);
                    xpl = (xp = nd_[x].parent) < 0 ? -1 : nd_[xp].left;
                }
                if (xpl < 0) x = xp;
                else {
                    int sl = nd_[xpl].left, sr = nd_[xpl].right;
                    if (!isRed(sl) && !isRed(sr)) { nd_[xpl].red = true; x = xp; }
                    else {
                        if (!isRed(sl)) {
                            if (sr >= 0) nd_[sr].red = false;
                            nd_[xpl].red = true;
                            root = rotateLeft(root, xpl);
                            xpl = (xp = nd_[x].parent) < 0 ? -1 : nd_[xp].left;
                        }
                        if (xpl >= 0) { nd_[xpl].red = xp < 0 ? false : nd_[xp].red; if ((sl = nd_[xpl].left) >= 0) nd_[sl].red = false; }
                        if (xp >= 0) { nd_[xp].red = false; root = rotateRight(root, xp); }
                        x = root;
                    }
                }
            }
        }
    }
    void removeTreeNode(int p, bool movable) {
        const size_t index = (uint32_t)nd_[p].hash & (tab_.size() - 1);
        int first = tab_[index], root = first, rl;
        const int succ = nd_[p].next, pred = nd_[p].prev;
        if (pred < 0) tab_[index] = first = succ; else nd_[pred].next = succ;
        if (succ >= 0) nd_[succ].prev = pred;
        if (first < 0) return;
        if (nd_[root].parent >= 0) root = rootOf(root);
        if (movable && (nd_[root].right < 0 || (rl = nd_[root].left) < 0 || nd_[rl].left < 0)) { untreeify(first); return; }   // too small
        const int pl = nd_[p].left, pr = nd_[p].right;
        int replacement;
        if (pl >= 0 && pr >= 0) {
            int s = pr, sl;
            while ((sl = nd_[s].left) >= 0) s = sl;   // successor
            std::swap(nd_[s].red, nd_[p].red);
            const int sr = nd_[s].right, pp = nd_[p].parent;
            if (s == pr) { nd_[p].parent = s; nd_[s].right = p; }
            else {
                const int sp = nd_[s].parent;
                if ((nd_[p].parent = sp) >= 0) { if (s == nd_[sp].left) nd_[sp].left = p; else nd_[sp].right = p; }
                if ((nd_[s].right = pr) >= 0) nd_[pr].parent = s;
            }
            nd_[p].left = -1;
            if ((nd_[p].right = sr) >= 0) nd_[sr].parent = p;
            if ((nd_[s].left = pl) >= 0) nd_[pl].parent = s;
            if ((nd_[s].parent = pp) < 0) root = s;
            else if (p == nd_[pp].left) nd_[pp].left = s;
            else nd_[pp].right = s;
            replacement = sr >= 0 ? sr : p;
        } else if (pl >= 0) replacement = pl;
        else if (pr >= 0) replacement = pr;
        else replacement = p;
        if (replacement != p) {
            const int pp = nd_[replacement].parent = nd_[p].parent;
            if (pp < 0) { root = replacement; nd_[root].red = false; }
            else if (p == nd_[pp].left) nd_[pp].left = replacement;
            else nd_[pp].right = replacement;
            nd_[p].left = nd_[p].right = nd_[p].parent = -1;
        }
        const int r = nd_[p].red ? root : balanceDeletion(root, replacement);
        if (replacement == p) {   // detach
            const int pp = nd_[p].parent;
            nd_[p].parent = -1;
            if (pp >= 0) { if (p == nd_[pp].left) nd_[pp].left = -1; else if (p == nd_[pp].right) nd_[pp].right = -1; }
        }
        if (movable) moveRootToFront(r);
    }
    // TreeNode.split(): relink into lo and hi lists preserving order, then untreeify (<= 6) or treeify again
    void split(int b, std::vector<int> &ntab, size_t index, size_t bit) {
        int loHead = -1, loTail = -1, hiHead = -1, hiTail = -1, lc = 0, hc = 0;
        for (int e = b, next; e >= 0; e = next) {
            next = nd_[e].next; nd_[e].next = -1;
            if (((uint32_t)nd_[e].hash & bit) == 0) { if ((nd_[e].prev = loTail) < 0) loHead = e; else nd_[loTail].next = e; loTail = e; ++lc; }
            else { if ((nd_[e].prev = hiTail) < 0) hiHead = e; else nd_[hiTail].next = e; hiTail = e; ++hc; }
        }
        tab_.swap(ntab);   // treeify()/moveRootToFront() work on the new table
        if (loHead >= 0) {
            if (lc <= UNTREEIFY_THRESHOLD) { untreeify(loHead); tab_[index] = loHead; }
            else { tab_[index] = loHead; if (hiHead >= 0) treeify(loHead); }   // else: is already treeified
        }
        if (hiHead >= 0) {
            if (hc <= UNTREEIFY_THRESHOLD) { untreeify(hiHead); tab_[index + bit] = hiHead; }
            else { tab_[index + bit] = hiHead; if (loHead >= 0) treeify(hiHead); }
        }
        tab_.swap(ntab);
    }
    void resize() {
        const size_t ocap = tab_.size();
        std::vector<int> nt(ocap * 2, -1);
        for (size_t j = 0; j < ocap; ++j) {
            const int e = tab_[j];
            if (e < 0) continue;
            tab_[j] = -1;
            if (nd_[e].next < 0) nt[(uint32_t)nd_[e].hash & (ocap * 2 - 1)] = e;
            else if (nd_[e].tree) split(e, nt, j, ocap);
            else {   // lo/hi split keeps the relative order
                int loHead = -1, loTail = -1, hiHead = -1, hiTail = -1;
                for (int q = e, next; q >= 0; q = next) {
                    next = nd_[q].next; nd_[q].next = -1;
                    if (((uint32_t)nd_[q].hash & ocap) == 0) { if (loTail < 0) loHead = q; else nd_[loTail].next = q; loTail = q; }
                    else { if (hiTail < 0) hiHead = q; else nd_[hiTail].next = q; hiTail = q; }
                }
                nt[j] = loHead; nt[j + ocap] = hiHead;
            }
        }
        tab_.swap(nt);
        threshold_ *= 2;
    }
    bool checkTree(int t, int blacks, int *bh, size_t *cnt) const {
        if (t < 0) { if (*bh < 0) *bh = blacks; return *bh == blacks; }
        const Node &n = nd_[t];
        ++*cnt;
        if (n.red && (isRed(n.left) || isRed(n.right))) return false;
        if (n.left >= 0 && (nd_[n.left].parent != t || dirOf(n.left, t) > 0)) return false;
        if (n.right >= 0 && (nd_[n.right].parent != t || dirOf(n.right, t) <= 0)) return false;
        const int b2 = blacks + (n.red ? 0 : 1);
        return checkTree(n.left, b2, bh, cnt) && checkTree(n.right, b2, bh, cnt);
    }
    std::vector<std::string> keys_;
    std::vector<Node> nd_;
    std::unordered_map<std::string, int> index_;
    std::vector<int> tab_;
    size_t size_ = 0, threshold_ = 0;
};

// String.format("%<width>.<prec>f", x): Java rounds the shortest round-trip decimal of the double HALF_UP,
// unlike C printf (exact binary value, half-even).  DistanceMatrixProcessor.java:82 (%8.6f),
// GtoRepGenomeProcessor.java (%8.4f), ClassifyProcessor.java:117 (%12.8f).
inline std::string format_fixed(double x, int width, int prec) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);   // shortest repr, fixed notation
    std::string s(buf, r.ptr);
    bool neg = !s.empty() && s[0] == '-';
    if (neg) s.erase(0, 1);
    size_t dot = s.find('.');
    std::string ip = dot == std::string::npos ? s : s.substr(0, dot);
    std::string fp = dot == std::string::npos ? "" : s.substr(dot + 1);
    bool up = (int)fp.size() > prec && fp[prec] >= '5';
    fp.resize(prec, '0');
    std::string digits = ip + fp;
    if (up) {
        int i = (int)digits.size() - 1;
        while (i >= 0 && digits[i] == '9') digits[i--] = '0';
        if (i >= 0) digits[i]++; else digits.insert(digits.begin(), '1');
    }
    size_t il = digits.size() - prec;
    std::string out = digits.substr(0, il);
    if (prec > 0) out += "." + digits.substr(il);
    if (neg && out.find_first_not_of("0.") != std::string::npos) out = "-" + out;
    if ((int)out.size() < width) out = std::string(width - out.size(), ' ') + out;
    return out;
}

// Double.toString(x) (NeighborProcessor.java:225-226): shortest repr; plain for 1e-3 <= |x| < 1e7, else d.dddE[-]n
inline std::string double_to_string(double x) {
    if (x == 0) return std::signbit(x) ? "-0.0" : "0.0";
    double ax = x < 0 ? -x : x;
    char buf[64];
    if (ax >= 1e-3 && ax < 1e7) {
        auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);
        std::string s(buf, r.ptr);
        if (s.find('.') == std::string::npos) s += ".0";
        return s;
    }
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
    std::string s(buf, r.ptr);   // d.ddde+XX
    size_t e = s.find('e');
    std::string mant = s.substr(0, e);
    int ex = std::atoi(s.c_str() + e + 1);
    if (mant.find('.') == std::string::npos) mant += ".0";
    return mant + "E" + std::to_string(ex);
}

}  // namespace java

// ---------------------------------------------------------------------------------------------------
struct Sequence {   // org.theseed.sequence.Sequence: label, comment, sequence
    std::string label, comment, sequence;
    const std::string &getLabel() const { return label; }
    const std::string &getComment() const { return comment; }
    const std::string &getSequence() const { return sequence; }
};

// Feature.genomeOf(fid) / RepGenome.getGenomeId(): the "digits.digits" of fig|G.peg.N
inline std::string genomeOf(const std::string &fid) {
    static const std::regex re("fig\\|(\\d+\\.\\d+)\\.");
    std::smatch m;
    if (std::regex_search(fid, m, re) && m.position(0) == 0) return m[1];
    return fid;
}

class FastaInputStream {   // iterates Sequence records; a header with no sequence lines gives an empty sequence
  public:
    explicit FastaInputStream(std::istream &in) : in_(&in) {}
    explicit FastaInputStream(const std::string &file) : owned_(new std::ifstream(file)), in_(owned_.get()) {
        if (!*owned_) throw std::runtime_error("java.io.FileNotFoundException: " + file);
    }
    bool next(Sequence &out) {
        std::string line;
        if (!have_header_) {
            while (std::getline(*in_, line)) { chomp(line); if (!line.empty() && line[0] == '>') { header_ = line; have_header_ = true; break; } }
            if (!have_header_) return false;
        }
        out = Sequence();
        size_t sp = header_.find_first_of(" \t", 1);
        out.label = header_.substr(1, sp == std::string::npos ? std::string::npos : sp - 1);
        if (sp != std::string::npos) { size_t b = header_.find_first_not_of(" \t", sp); if (b != std::string::npos) out.comment = trim(header_.substr(b)); }
        have_header_ = false;
        while (std::getline(*in_, line)) {
            chomp(line);
            if (!line.empty() && line[0] == '>') { header_ = line; have_header_ = true; break; }
            out.sequence += trim(line);
        }
        return true;
    }
    std::vector<Sequence> readAll() { std::vector<Sequence> v; Sequence s; while (next(s)) v.push_back(s); return v; }
  private:
    static void chomp(std::string &s) { while (!s.empty() && (s.back() == '\r' || s.back() == '\n')) s.pop_back(); }
    static std::string trim(const std::string &s) {
        size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
        return a == std::string::npos ? "" : s.substr(a, b - a + 1);
    }
    std::unique_ptr<std::ifstream> owned_;
    std::istream *in_;
    std::string header_;
    bool have_header_ = false;
};

class FastaOutputStream {   // ">label comment", sequence wrapped at 60, LF (format of data/rep50.db)
  public:
    explicit FastaOutputStream(std::ostream &out) : out_(&out) {}
    void write(const Sequence &s) { write(s.label, s.comment, s.sequence); }
    void write(const std::string &label, const std::string &comment, const std::string &seq) {
        *out_ << '>' << label;
        if (!comment.empty()) *out_ << ' ' << comment;
        *out_ << '\n';
        for (size_t i = 0; i < seq.size(); i += 60) *out_ << seq.substr(i, 60) << '\n';
    }
  private:
    std::ostream *out_;
};

// ---------------------------------------------------------------------------------------------------
// The GPU engine behind the classes: one kr_ctx per k-mer size on device $KR_DEVICE (default 0).
// With KR_DEVICES=0,1,... (a device may be listed more than once) there is one kr_multi per k-mer size instead: the batch
// calls of the classes below shard over all listed GPUs inside the library, single-object calls use its first context.
class KmerEngine {
  public:
    static kr_multi *multi(int key) {
        const char *list = std::getenv("KR_DEVICES");
        if (!list || !*list) return nullptr;
        auto &m = multiMap();
        auto it = m.find(key);
        if (it != m.end()) return it->second;
        std::vector<int> devs;
        for (const char *p = list; *p;) { devs.push_back(std::atoi(p)); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
        uint32_t flags = KR_IDENT_IS_INFINITY;
        if (const char *f = std::getenv("KR_FLAGS")) flags = (uint32_t)std::atoi(f);
        kr_multi *h = nullptr;
        if (kr_multi_create(devs.data(), (int)devs.size(), key, flags, &h) != KR_OK)
            throw std::runtime_error(std::string("kmersreps: ") + kr_multi_last_error(nullptr));
        m[key] = h;
        return h;
    }
    static void checkMulti(kr_multi *m, kr_status st) { if (st != KR_OK) throw std::runtime_error(std::string("kmersreps: ") + kr_multi_last_error(m)); }
    static kr_ctx *ctx(int key) {
        auto &m = map();
        auto it = m.find(key);
        if (it != m.end()) return it->second;
        if (kr_multi *mh = multi(key)) return m[key] = kr_multi_ctx(mh, 0);
        const char *d = std::getenv("KR_DEVICE");
        uint32_t flags = KR_IDENT_IS_INFINITY;
        if (const char *f = std::getenv("KR_FLAGS")) flags = (uint32_t)std::atoi(f);
        kr_ctx *c = nullptr;
        if (kr_create(d ? std::atoi(d) : 0, key, flags, &c) != KR_OK)
            throw std::runtime_error(std::string("kmersreps: ") + kr_last_error(nullptr));   // RuntimeException, App.java:129
        m[key] = c;
        return c;
    }
    static void check(kr_ctx *c, kr_status st) { if (st != KR_OK) throw std::runtime_error(std::string("kmersreps: ") + kr_last_error(c)); }
  private:
    static std::map<int, kr_ctx *> &map() { static std::map<int, kr_ctx *> m; return m; }
    static std::map<int, kr_multi *> &multiMap() { static std::map<int, kr_multi *> m; return m; }
};

// flat residue buffer + offsets of a list of strings (what every batch entry point takes)
struct FlatSeqs {
    std::string res; std::vector<uint64_t> off{0};
    void add(const std::string &s) { res += s; off.push_back(res.size()); }
    const uint8_t *data() const { static const uint8_t none = 0; return res.empty() ? &none : (const uint8_t *)res.data(); }
    uint32_t count() const { return (uint32_t)off.size() - 1; }
};

struct SetHandle {   // shared ownership of an encoded batch
    kr_ctx *ctx; kr_seqset *set; int K;
    SetHandle(kr_ctx *c, kr_seqset *s, int k) : ctx(c), set(s), K(k) {}
    ~SetHandle() { kr_seqset_free(set); }
};

inline std::shared_ptr<SetHandle> encodeStrings(int key, const std::vector<const std::string *> &seqs) {
    std::string flat; std::vector<uint64_t> off(seqs.size() + 1, 0);
    for (size_t i = 0; i < seqs.size(); ++i) { flat += *seqs[i]; off[i + 1] = flat.size(); }
    kr_ctx *c = KmerEngine::ctx(key);
    kr_seqset *s = nullptr;
    static const uint8_t none = 0;
    KmerEngine::check(c, kr_encode(c, flat.empty() ? &none : (const uint8_t *)flat.data(), off.data(), (uint32_t)seqs.size(), &s));
    return std::make_shared<SetHandle>(c, s, key);
}

// ---------------------------------------------------------------------------------------------------
// org.theseed.sequence.SequenceKmers: the abstract base of ProteinKmers -- the k-mer set of one sequence,
// similarity / distance / size / INFINITY.  The set itself lives on the GPU (a batch handle + an index into it).
class SequenceKmers {
  public:
    static constexpr int INFINITY_ = KR_INFINITY;   // SequenceKmers.INFINITY (RepGenomeDbProcessor.java:249)
    virtual ~SequenceKmers() = default;
    SequenceKmers() = default;
    const std::string &getSequence() const { return sequence_; }
    int getLen() const { return (int)sequence_.size(); }   // ProteinDataFactory.java:308

    // number of distinct k-mers (ScoreCouplingReporter.java:73)
    int size() const { ensure(); return (int)size_; }
    // number of shared distinct k-mers, INFINITY for equal sequences (DistanceMatrixProcessor.java:76)
    int similarity(const SequenceKmers &other) const {
        ensure(); other.ensure();
        int32_t sim = 0; uint32_t ia = idx_, ib = other.idx_;
        KmerEngine::check(set_->ctx, kr_sim_pairs(set_->ctx, set_->set, &ia, other.set_->set, &ib, 1, &sim, nullptr));
        return sim;
    }
    // Jaccard distance (SequenceInfo.java:75, DistanceMatrixProcessor.java:75)
    double distance(const SequenceKmers &other) const { return distance(similarity(other), *this, other); }
    static double distance(int sim, const SequenceKmers &a, const SequenceKmers &b) {   // SsuRepresentation.java:47
        return kr_distance(sim, (uint32_t)a.size(), (uint32_t)b.size());
    }
    // the k-mers themselves (SequenceKmers.iterator(), TargetProcessor.java:193)
    std::vector<std::string> kmers() const {
        ensure();
        std::vector<uint64_t> keys(size_ ? size_ : 1); uint32_t n = 0;
        KmerEngine::check(set_->ctx, kr_seqset_keys(set_->set, idx_, keys.data(), (uint32_t)keys.size(), &n));
        std::vector<std::string> out(n, std::string((size_t)k_, ' '));
        for (uint32_t i = 0; i < n; ++i) kr_key_decode(set_->ctx, keys[i], out[i].data());
        return out;
    }

    // the sorted packed k-mers (kr_seqset_keys)
    std::vector<uint64_t> keys() const {
        ensure();
        std::vector<uint64_t> k(size_ ? size_ : 1); uint32_t n = 0;
        KmerEngine::check(set_->ctx, kr_seqset_keys(set_->set, idx_, k.data(), (uint32_t)k.size(), &n));
        k.resize(n);
        return k;
    }
    // retainAll(other) / removeAll(other) (TargetProcessor.java:166,184): the merge of the two sorted key lists becomes
    // this object's set (kr_seqset_from_keys; the sequence itself is unchanged, as upstream)
    void retainAll(const SequenceKmers &other) { edit(other, true); }
    void removeAll(const SequenceKmers &other) { edit(other, false); }
    // same sequence and same k-mers (SsuGenome.java:77-104 reaches these through super)
    bool equals(const SequenceKmers &o) const { return sequence_ == o.sequence_ && ((!edited_ && !o.edited_) || keys() == o.keys()); }
    size_t hashCode() const { size_t h = std::hash<std::string>()(sequence_); if (edited_) for (uint64_t k : keys()) h = h * 31 + (size_t)k; return h; }
    bool isEdited() const { return edited_; }

    // Encode many objects in one launch (the constructors never touch the device: they are called from parallel
    // streams, GtoRepBuildProcessor.java:127).  Objects already encoded are left alone.
    template <class It> static void encodeBatch(It first, It last) {
        std::vector<const SequenceKmers *> todo; std::vector<const std::string *> seqs;
        for (It it = first; it != last; ++it) { const SequenceKmers &p = deref(*it); if (!p.set_) { todo.push_back(&p); seqs.push_back(&p.sequence_); } }
        if (todo.empty()) return;
        auto h = encodeStrings(todo[0]->k(), seqs);
        std::vector<uint32_t> sizes(todo.size());
        KmerEngine::check(h->ctx, kr_seqset_sizes(h->set, sizes.data()));
        for (size_t i = 0; i < todo.size(); ++i) { todo[i]->set_ = h; todo[i]->idx_ = (uint32_t)i; todo[i]->size_ = sizes[i]; }
    }
    const std::shared_ptr<SetHandle> &handle() const { ensure(); return set_; }
    uint32_t index() const { ensure(); return idx_; }
    int k() const { return k_; }   // engine key: the k-mer size

  protected:
    SequenceKmers(std::string seq, int k) : sequence_(std::move(seq)), k_(k) {}
    static const SequenceKmers &deref(const SequenceKmers &p) { return p; }
    static const SequenceKmers &deref(const SequenceKmers *p) { return *p; }
    template <class T> static const SequenceKmers &deref(const std::unique_ptr<T> &p) { return *p; }
    template <class T> static const SequenceKmers &deref(const std::shared_ptr<T> &p) { return *p; }
    void ensure() const { if (!set_) { const SequenceKmers *self = this; encodeBatch(&self, &self + 1); } }
    void edit(const SequenceKmers &other, bool keepCommon) {
        const std::vector<uint64_t> a = keys(), b = other.keys();
        std::vector<uint64_t> out;
        size_t j = 0;
        for (uint64_t x : a) {
            while (j < b.size() && b[j] < x) ++j;
            const bool common = j < b.size() && b[j] == x;
            if (common == keepCommon) out.push_back(x);
        }
        kr_ctx *c = KmerEngine::ctx(k_);
        const uint64_t koff[2] = {0, out.size()}, roff[2] = {0, sequence_.size()};
        static const uint64_t nokey = 0; static const uint8_t none = 0;
        kr_seqset *s = nullptr;
        KmerEngine::check(c, kr_seqset_from_keys(c, out.empty() ? &nokey : out.data(), koff, sequence_.empty() ? &none : (const uint8_t *)sequence_.data(), roff, 1, &s));
        set_ = std::make_shared<SetHandle>(c, s, k_);
        idx_ = 0; size_ = (uint32_t)out.size(); edited_ = true;
    }
    std::string sequence_;
    int k_ = 8;
    mutable std::shared_ptr<SetHandle> set_;
    mutable uint32_t idx_ = 0, size_ = 0;
    bool edited_ = false;
};

class ProteinKmers : public SequenceKmers {
  public:
    static void setKmerSize(int k) { K() = k; }   // process-global, as in the reference (RepGenomeDbProcessor.java:77-80)
    static int kmerSize() { return K(); }
    ProteinKmers() = default;
    explicit ProteinKmers(std::string protein) : SequenceKmers(std::move(protein), K()) {}   // SequenceInfo.java:47
    const std::string &getProtein() const { return sequence_; }
  protected:
    static int &K() { static int k = 8; return k; }
};

// All-vs-all similarities of a list of ProteinKmers (the rewritten double loops of
// DistanceMatrixProcessor.java:71-76 and RepMatrixProcessor.java:88-93).
class SimilarityMatrix {
  public:
    // The matrix is fetched as the cells with similarity >= 1 of the strict upper triangle (kr_sim_matrix_sparse: a few
    // cells per sequence cross PCIe instead of N * N values) plus the diagonal; every pair not listed has similarity 0 by
    // definition.  The callers' loops still visit every pair in the reference's order (DistanceMatrixProcessor.java:71-83,
    // SequenceInfo.java:83-89) and read zeros here, so their output and their floating-point sums are unchanged.
    template <class It> static SimilarityMatrix compute(It first, It last) {
        SequenceKmers::encodeBatch(first, last);
        SimilarityMatrix m;
        std::vector<const SequenceKmers *> v;
        for (It it = first; it != last; ++it) v.push_back(&ref(*it));
        m.n_ = v.size();
        if (!m.n_) return m;
        bool edited = false;
        for (auto p : v) edited = edited || p->isEdited();
        std::vector<uint32_t> cells;       // kr_multi path: 16-byte cells
        std::vector<uint64_t> cells64;     // single context: packed 8-byte cells (half the bytes over PCIe)
        uint64_t count = 0, cap = std::max<uint64_t>(4096, 16 * (uint64_t)m.n_);
        m.sizes_.resize(m.n_);
        m.diag_.resize(m.n_);
        std::shared_ptr<SetHandle> h;
        kr_multi *mh = edited ? nullptr : KmerEngine::multi(v[0]->k());
        FlatSeqs f;
        if (mh) {      // every listed GPU scans a row stripe
            for (auto p : v) f.add(p->getSequence());
            for (size_t i = 0; i < m.n_; ++i) m.sizes_[i] = (uint32_t)v[i]->size();
        } else {       // one set holding exactly these objects (gathered / re-encoded if they come from different batches)
            h = v[0]->handle();
            bool same = true;
            for (size_t i = 0; i < v.size(); ++i) same = same && v[i]->handle() == h && v[i]->index() == i;
            if (!same || kr_seqset_count(h->set) != m.n_) {
                if (edited) throw std::runtime_error("SimilarityMatrix: objects edited by retainAll / removeAll must come from one batch");
                std::vector<const std::string *> s; for (auto p : v) s.push_back(&p->getSequence()); h = encodeStrings(v[0]->k(), s);
            }
            KmerEngine::check(h->ctx, kr_seqset_sizes(h->set, m.sizes_.data()));
        }
        for (;;) {     // the library reports how many cells qualify when the room was too small: once more with room for all
            kr_status st;
            const bool packed = !mh && m.n_ <= (1u << 24);
            if (mh) { cells.resize((size_t)cap * 4); st = kr_multi_sim_matrix_sparse(mh, f.data(), f.off.data(), f.count(), nullptr, nullptr, 0, 1, KR_LAYOUT_TRIANGLE, cells.data(), cap, &count); }
            else {
                kr_matrix_desc d{KR_ELEM_I32, KR_LAYOUT_TRIANGLE, 0, (uint32_t)m.n_};
                if (packed) { cells64.resize((size_t)cap); st = kr_sim_matrix_sparse64(h->ctx, h->set, nullptr, &d, 1, cells64.data(), cap, &count); }
                else { cells.resize((size_t)cap * 4); st = kr_sim_matrix_sparse(h->ctx, h->set, nullptr, &d, 1, cells.data(), cap, &count); }
            }
            if (st != KR_OK && count > cap) { cap = count; continue; }
            if (mh) KmerEngine::checkMulti(mh, st); else KmerEngine::check(h->ctx, st);
            if (packed) {      // unpack into the (row, col, sim, 0) form the CSR build below reads
                cells.resize((size_t)count * 4);
                for (uint64_t k = 0; k < count; ++k) { const uint64_t c = cells64[k]; cells[4 * k] = KR_CELL_ROW(c); cells[4 * k + 1] = KR_CELL_COL(c); cells[4 * k + 2] = (uint32_t)KR_CELL_SIM(c); }
            }
            break;
        }
        // an object and itself (INFINITY under the equal-sequence rule, else its size): one launch of explicit pairs
        {
            std::vector<uint32_t> idx(m.n_);
            for (size_t i = 0; i < m.n_; ++i) idx[i] = (uint32_t)i;
            if (mh) { for (size_t i = 0; i < m.n_; ++i) m.diag_[i] = v[i]->similarity(*v[i]); }
            else KmerEngine::check(h->ctx, kr_sim_pairs(h->ctx, h->set, idx.data(), h->set, idx.data(), m.n_, m.diag_.data(), nullptr));
        }
        // CSR over the rows of the strict upper triangle, columns ascending
        m.start_.assign(m.n_ + 1, 0);
        for (uint64_t k = 0; k < count; ++k) m.start_[cells[4 * k] + 1]++;
        for (size_t i = 0; i < m.n_; ++i) m.start_[i + 1] += m.start_[i];
        m.col_.resize(count); m.val_.resize(count);
        std::vector<uint64_t> at(m.start_.begin(), m.start_.end() - 1);
        for (uint64_t k = 0; k < count; ++k) { const uint64_t p = at[cells[4 * k]]++; m.col_[p] = cells[4 * k + 1]; m.val_[p] = (int32_t)cells[4 * k + 2]; }
        for (size_t i = 0; i < m.n_; ++i) {      // cells arrive in no particular order: sort each row by column
            const uint64_t a = m.start_[i], b = m.start_[i + 1];
            if (b - a < 2) continue;
            std::vector<std::pair<uint32_t, int32_t>> row;
            for (uint64_t p = a; p < b; ++p) row.emplace_back(m.col_[p], m.val_[p]);
            std::sort(row.begin(), row.end());
            for (uint64_t p = a; p < b; ++p) { m.col_[p] = row[p - a].first; m.val_[p] = row[p - a].second; }
        }
        return m;
    }
    int similarity(size_t i, size_t j) const {
        if (i == j) return diag_[i];
        if (i > j) std::swap(i, j);
        const uint32_t *lo = col_.data() + start_[i], *hi = col_.data() + start_[i + 1];
        const uint32_t *p = std::lower_bound(lo, hi, (uint32_t)j);
        return (p != hi && *p == j) ? val_[p - col_.data()] : 0;
    }
    double distance(size_t i, size_t j) const { return kr_distance(similarity(i, j), sizes_[i], sizes_[j]); }
    size_t size() const { return n_; }
    uint64_t cells() const { return col_.size(); }      // pairs that share at least one k-mer
  private:
    static const SequenceKmers &ref(const SequenceKmers &p) { return p; }
    static const SequenceKmers &ref(const SequenceKmers *p) { return *p; }
    template <class T> static const SequenceKmers &ref(const std::unique_ptr<T> &p) { return *p; }
    size_t n_ = 0;
    std::vector<uint64_t> start_;
    std::vector<uint32_t> col_;
    std::vector<int32_t> val_, diag_;
    std::vector<uint32_t> sizes_;
};

// ---------------------------------------------------------------------------------------------------
class RepGenome : public ProteinKmers {   // a representative: (fid, name, protein)
  public:
    RepGenome() = default;
    explicit RepGenome(const Sequence &seq) : RepGenome(seq.label, seq.comment, seq.sequence) {}   // RepGenomeDbProcessor.java:161
    RepGenome(std::string fid, std::string name, std::string protein)                               // BaseGenomeProcessor.java:201-202
        : ProteinKmers(std::move(protein)), fid_(std::move(fid)), name_(std::move(name)), genomeId_(genomeOf(fid_)) {}
    const std::string &getFid() const { return fid_; }
    const std::string &getName() const { return name_; }
    const std::string &getGenomeId() const { return genomeId_; }
    // Used only to visit each unordered pair once (DistanceMatrixProcessor.java:74) and to break ties in sorted reports.
    // The class is in un-vendored org.theseed:shared, so which order its compareTo imposes on genome ids cannot be read
    // here (open semantic questions): a run-time policy, not a guess baked in.  LEXICAL = String.compareTo
    // of the ids (default); NATURAL = taxon number, then version number ("83333.1" < "511145.12").  KR_ID_ORDER=natural
    // in the environment (or setIdOrder) selects the other; values never change, only row order / pair orientation.
    enum class IdOrder { LEXICAL, NATURAL };
    static IdOrder &idOrder() {
        static IdOrder o = [] { const char *e = std::getenv("KR_ID_ORDER"); return (e && std::string(e) == "natural") ? IdOrder::NATURAL : IdOrder::LEXICAL; }();
        return o;
    }
    static void setIdOrder(IdOrder o) { idOrder() = o; }
    static int compareIds(const std::string &a, const std::string &b) {
        if (idOrder() == IdOrder::NATURAL) {
            // dot-separated fields, numeric where both are digit strings (no leading-zero ids exist; longer = larger)
            size_t i = 0, j = 0;
            while (i < a.size() && j < b.size()) {
                size_t ie = a.find('.', i), je = b.find('.', j);
                if (ie == std::string::npos) ie = a.size();
                if (je == std::string::npos) je = b.size();
                const std::string fa = a.substr(i, ie - i), fb = b.substr(j, je - j);
                const bool na = !fa.empty() && fa.find_first_not_of("0123456789") == std::string::npos;
                const bool nb = !fb.empty() && fb.find_first_not_of("0123456789") == std::string::npos;
                int c;
                if (na && nb) {
                    const size_t za = std::min(fa.find_first_not_of('0'), fa.size() - 1), zb = std::min(fb.find_first_not_of('0'), fb.size() - 1);
                    const size_t la = fa.size() - za, lb = fb.size() - zb;
                    c = la != lb ? (la < lb ? -1 : 1) : fa.compare(za, la, fb, zb, lb);
                } else c = fa.compare(fb);
                if (c) return c < 0 ? -1 : 1;
                i = ie + 1; j = je + 1;
            }
            const bool ra = i < a.size(), rb = j < b.size();
            if (ra != rb) return ra ? 1 : -1;
            // numerically equal but differently written ids stay distinct: fall through to the string order
        }
        const int c = a.compare(b);
        return c < 0 ? -1 : c > 0 ? 1 : 0;
    }
    int compareTo(const RepGenome &o) const { return compareIds(genomeId_, o.genomeId_); }
    std::string toString() const { return genomeId_ + " (" + name_ + ")"; }
  private:
    std::string fid_, name_, genomeId_;
};

class RepGenomeDb {
  public:
    static constexpr const char *DEFAULT_PROTEIN = "Phenylalanyl-tRNA synthetase alpha chain";   // RepGenomeDbProcessor.java:184

    // RepGenomeDb.Representation (RepGenomeDbProcessor.java:237-252, GtoRepGenomeProcessor.java:271-292)
    class Representation {
      public:
        Representation() = default;
        Representation(const RepGenomeDb *db, const RepGenome *rep, int sim, double dist) : db_(db), rep_(rep), sim_(sim), dist_(dist) {}
        // Representation(String repId, int score, double distance): a stored result without the object (ProteinData.java:242)
        Representation(const RepGenomeDb *db, std::string repId, int score, double dist) : db_(db), sim_(score), dist_(dist), repId_(std::move(repId)) {}
        int getSimilarity() const { return sim_; }
        double getDistance() const { return dist_; }
        const RepGenome *getRepresentative() const { return rep_; }
        std::string getGenomeId() const { return rep_ ? rep_->getGenomeId() : repId_; }
        bool isRepresented() const { return db_ && sim_ >= db_->getThreshold(); }
      private:
        const RepGenomeDb *db_ = nullptr; const RepGenome *rep_ = nullptr; int sim_ = 0; double dist_ = 1.0; std::string repId_;
    };

    explicit RepGenomeDb(int threshold, std::string protName = DEFAULT_PROTEIN)   // RepGenomeDbProcessor.java:122
        : threshold_(threshold), kmerSize_(ProteinKmers::kmerSize()), protName_(std::move(protName)) {}
    ~RepGenomeDb() { if (db_) kr_db_free(db_); if (mdb_) kr_multi_db_free(mdb_); }
    RepGenomeDb(const RepGenomeDb &) = delete;
    RepGenomeDb &operator=(const RepGenomeDb &) = delete;

    int getThreshold() const { return threshold_; }
    int getKmerSize() const { return kmerSize_; }
    const std::string &getProtName() const { return protName_; }
    size_t size() const { return order_.size(); }
    bool contains(const std::string &genomeId) const { return order_.contains(genomeId); }
    const RepGenome *get(const std::string &genomeId) const { int i = order_.find(genomeId); return i < 0 ? nullptr : reps_[i].get(); }
    std::string getListFileName() const { return "rep" + std::to_string(threshold_) + ".list.tbl"; }   // BaseGenomeProcessor.java:357
    std::string toString() const { return "RepGen" + std::to_string(threshold_); }

    // iteration in java.util.HashMap order by genome id
    std::vector<const RepGenome *> all() const { std::vector<const RepGenome *> v; for (int i : order_.order()) v.push_back(reps_[i].get()); return v; }
    std::vector<const RepGenome *> sorted() const {   // SortedSet<RepGenome> (RepGenomeDbProcessor.java:214)
        auto v = all();
        std::sort(v.begin(), v.end(), [](const RepGenome *a, const RepGenome *b) { return a->compareTo(*b) < 0; });
        return v;
    }

    // put(genomeId, rep): a second rep for the same genome replaces the first in place (CopyTest.java:16-27)
    void addRep(const RepGenome &rep) {
        bool is_new = false;
        int id = order_.put(rep.getGenomeId(), &is_new);
        if (is_new) reps_.push_back(std::make_unique<RepGenome>(rep));
        else { *reps_[id] = rep; rebuild_ = true; }
        ranks_dirty_ = true;
    }
    void remove(const std::string &genomeId) {   // iterator().remove() (UpdateProcessor.java:96-103)
        if (order_.contains(genomeId)) { order_.remove(genomeId); rebuild_ = true; ranks_dirty_ = true; }
    }

    // findClosest(ProteinKmers | Sequence | RepGenome): best similarity over all reps, first strict maximum
    // in iteration order; (0, null, 1.0) when nothing shares a k-mer (GtoRepGenomeProcessor.java:277-282)
    Representation findClosest(const SequenceKmers &q) const { const SequenceKmers *p = &q; return findClosest(&p, &p + 1)[0]; }
    Representation findClosest(const Sequence &seq) const { ProteinKmers q(seq.sequence); return findClosest(q); }   // RepGenomeDbProcessor.java:237
    template <class It> std::vector<Representation> findClosest(It first, It last) const {
        std::vector<Representation> out;
        SequenceKmers::encodeBatch(first, last);
        std::vector<const SequenceKmers *> v;
        for (It it = first; it != last; ++it) v.push_back(ptr(*it));
        if (v.empty()) return out;
        sync();
        std::vector<int32_t> sim(v.size()); std::vector<uint32_t> f(v.size()), l(v.size());
        if (kr_multi *mh = multi()) {      // the queries are split over the GPUs inside the library
            FlatSeqs fq; for (auto p : v) fq.add(p->getSequence());
            KmerEngine::checkMulti(mh, kr_multi_db_find_closest(mh, mdb_, fq.data(), fq.off.data(), fq.count(), sim.data(), f.data(), l.data()));
        } else {
            auto h = batchOf(v);
            KmerEngine::check(ctx(), kr_db_find_closest(ctx(), db_, h->set, sim.data(), f.data(), l.data()));
        }
        for (size_t i = 0; i < v.size(); ++i) {
            uint32_t pick = tie_last_ ? l[i] : f[i];
            if (sim[i] == 0 || pick == KR_NO_REP) { out.emplace_back(this, nullptr, 0, 1.0); continue; }
            const RepGenome *rep = reps_[dev2host_[pick]].get();
            out.emplace_back(this, rep, sim[i], SequenceKmers::distance(sim[i], *v[i], *rep));
        }
        return out;
    }

    // Similarity of every query to EVERY representative: row q holds `new Representation(rep, query).getSimilarity()`
    // for the reps in iteration order (the rewritten `for (RepGenome rep : this.reps)` loop of
    // NeighborProcessor.java:165-168).  reps_out receives that order.  One rectangular GPU matrix per query chunk.
    template <class It> std::vector<int32_t> similarities(It first, It last, std::vector<const RepGenome *> *reps_out) const {
        std::vector<const RepGenome *> reps = all();
        if (reps_out) *reps_out = reps;
        SequenceKmers::encodeBatch(first, last);
        std::vector<const SequenceKmers *> v;
        for (It it = first; it != last; ++it) v.push_back(ptr(*it));
        std::vector<int32_t> out(v.size() * reps.size());
        if (v.empty() || reps.empty()) return out;
        SequenceKmers::encodeBatch(reps.begin(), reps.end());   // one launch; the reps keep their sizes for distance()
        if (kr_multi *mh = multi()) {      // query rows split over the GPUs
            FlatSeqs fq, fr;
            for (auto p : v) fq.add(p->getSequence());
            for (auto r : reps) fr.add(r->getSequence());
            KmerEngine::checkMulti(mh, kr_multi_sim_matrix(mh, fq.data(), fq.off.data(), fq.count(), fr.data(), fr.off.data(), fr.count(), KR_ELEM_I32, KR_LAYOUT_FULL, out.data()));
            return out;
        }
        auto hr = batchOf(std::vector<const SequenceKmers *>(reps.begin(), reps.end()));
        const size_t R = reps.size(), chunk = std::max<size_t>(256, ((size_t)1 << 26) / R / 256 * 256);   // <= 64M cells per launch
        for (size_t q0 = 0; q0 < v.size(); q0 += chunk) {
            const size_t n = std::min(chunk, v.size() - q0);
            std::vector<const SequenceKmers *> part(v.begin() + q0, v.begin() + q0 + n);
            auto hq = batchOf(part);
            kr_matrix_desc d{KR_ELEM_I32, KR_LAYOUT_FULL, 0, (uint32_t)n};
            KmerEngine::check(hq->ctx, kr_sim_matrix(hq->ctx, hq->set, hr->set, &d, out.data() + q0 * R));
        }
        return out;
    }
    // The N closest representatives of one query, with the reference's tie retention (NeighborProcessor.java:161-193):
    // an insertion-sorted list, descending by similarity; a new result goes BEFORE the entries it ties with; when the
    // list outgrows `num` it is cut back to `num` only if the entry at the break beats the one after it, so ties
    // across the break are all kept.  `row` = similarities in iteration order; returns positions into it.
    static std::vector<size_t> closestN(const int32_t *row, size_t R, int num, int minSim) {
        std::vector<size_t> found;
        for (size_t r = 0; r < R; ++r) {
            const int sim = row[r];
            if (sim < minSim) continue;
            size_t idx = 0; const size_t n = found.size();
            while (idx < n && row[found[idx]] > sim) idx++;
            if (idx < n || n < (size_t)num) {
                found.insert(found.begin() + idx, r);
                if (found.size() > (size_t)num && row[found[num - 1]] > row[found[num]]) found.resize(num);
            }
        }
        return found;
    }

    // checkSimilarity(ProteinKmers, sim): is some rep at least this similar? (BaseGenomeProcessor.java:203)
    bool checkSimilarity(const SequenceKmers &q, int sim) const {
        if (size() == 0) return false;
        sync();
        uint8_t r = 0;
        if (kr_multi *mh = multi()) {
            FlatSeqs fq; fq.add(q.getSequence());
            KmerEngine::checkMulti(mh, kr_multi_db_any_ge(mh, mdb_, fq.data(), fq.off.data(), 1, sim, &r));
            return r != 0;
        }
        const SequenceKmers *p = &q; auto h = batchOf(std::vector<const SequenceKmers *>{p});
        KmerEngine::check(ctx(), kr_db_any_ge(ctx(), db_, h->set, sim, &r));
        return r != 0;
    }
    // checkGenome(rep): add it unless it is represented; returns true when added (RepGenomeDbProcessor.java:162)
    bool checkGenome(const RepGenome &rep) { std::vector<RepGenome> one{rep}; return checkGenomes(one)[0]; }
    // the same for a list, in list order, with the exact sequential semantics (batched on the GPU)
    std::vector<bool> checkGenomes(const std::vector<RepGenome> &cands, uint32_t batch = 0) {
        std::vector<bool> added(cands.size(), false);
        if (cands.empty()) return added;
        sync();
        std::vector<uint8_t> a(cands.size());
        if (kr_multi *mh = multi()) {      // scoring of every batch split over the GPUs, replicas kept identical
            FlatSeqs fc; for (auto &c : cands) fc.add(c.getSequence());
            KmerEngine::checkMulti(mh, kr_multi_greedy_build(mh, mdb_, fc.data(), fc.off.data(), fc.count(), threshold_, batch, a.data()));
        } else {
            SequenceKmers::encodeBatch(cands.begin(), cands.end());
            std::vector<const SequenceKmers *> v; for (auto &c : cands) v.push_back(&c);
            auto h = batchOf(v);
            KmerEngine::check(ctx(), kr_greedy_build(ctx(), db_, h->set, threshold_, batch, a.data()));
        }
        for (size_t i = 0; i < cands.size(); ++i)
            if (a[i]) {
                added[i] = true;
                bool is_new = false;
                int id = order_.put(cands[i].getGenomeId(), &is_new);
                if (is_new) { reps_.push_back(std::make_unique<RepGenome>(cands[i])); dev2host_.push_back(id); }
                else { *reps_[id] = cands[i]; rebuild_ = true; }   // duplicate genome id: replaced in place
            }
        synced_ = reps_.size();
        ranks_dirty_ = true;
        return added;
    }

    // save(File) / load(File): header ">Rep<thr>,K=<K> <protName>" with no sequence, then one record per rep
    // in iteration order (data/rep50.db:1-8; RepGenomeDbProcessor.java:152,209)
    void save(std::ostream &out) const {
        out << ">Rep" << threshold_ << ",K=" << kmerSize_ << ' ' << protName_ << '\n';
        FastaOutputStream fo(out);
        for (const RepGenome *r : all()) fo.write(r->getFid(), r->getName(), r->getProtein());
    }
    void save(const std::string &file) const {
        std::ofstream f(file, std::ios::binary);
        if (!f) throw std::runtime_error("java.io.IOException: cannot write " + file);
        save(f);
    }
    static std::unique_ptr<RepGenomeDb> load(std::istream &in) {
        FastaInputStream fi(in);
        Sequence hdr;
        if (!fi.next(hdr)) throw std::runtime_error("java.io.IOException: empty representative-genome file");
        static const std::regex re("Rep(\\d+),K=(\\d+)");
        std::smatch m;
        if (!std::regex_match(hdr.label, m, re)) throw std::runtime_error("java.io.IOException: invalid rep-genome header " + hdr.label);
        ProteinKmers::setKmerSize(std::stoi(m[2]));
        auto db = std::make_unique<RepGenomeDb>(std::stoi(m[1]), hdr.comment);
        Sequence s;
        while (fi.next(s)) db->addRep(RepGenome(s));
        return db;
    }
    static std::unique_ptr<RepGenomeDb> load(const std::string &file) {
        std::ifstream f(file, std::ios::binary);
        if (!f) throw std::runtime_error("java.io.FileNotFoundException: " + file);
        return load(f);
    }
    void setTieLast(bool v) { tie_last_ = v; }

  private:
    static const SequenceKmers *ptr(const SequenceKmers &p) { return &p; }
    static const SequenceKmers *ptr(const SequenceKmers *p) { return p; }
    kr_ctx *ctx() const { return KmerEngine::ctx(kmerSize_); }
    kr_multi *multi() const { return KmerEngine::multi(kmerSize_); }
    // the device copy: a kr_db on the single context, or a kr_multi_db (one replica per GPU) under KR_DEVICES
    void devFree() const { if (db_) { kr_db_free(db_); db_ = nullptr; } if (mdb_) { kr_multi_db_free(mdb_); mdb_ = nullptr; } }
    bool devExists() const { return db_ || mdb_; }
    void devCreate() const {
        if (kr_multi *mh = multi()) KmerEngine::checkMulti(mh, kr_multi_db_create(mh, &mdb_));
        else KmerEngine::check(ctx(), kr_db_create(ctx(), &db_));
    }
    void devAppend(const FlatSeqs &f) const {
        if (kr_multi *mh = multi()) KmerEngine::checkMulti(mh, kr_multi_db_append(mh, mdb_, f.data(), f.off.data(), f.count()));
        else KmerEngine::check(ctx(), kr_db_append(ctx(), db_, f.data(), f.off.data(), f.count()));
    }
    void devSetRanks(const std::vector<uint32_t> &rank) const {
        if (kr_multi *mh = multi()) KmerEngine::checkMulti(mh, kr_multi_db_set_ranks(mh, mdb_, rank.data(), (uint32_t)rank.size()));
        else KmerEngine::check(ctx(), kr_db_set_ranks(ctx(), db_, rank.data(), (uint32_t)rank.size()));
    }
    // one set holding exactly these objects, in order
    static std::shared_ptr<SetHandle> batchOf(const std::vector<const SequenceKmers *> &v) {
        std::shared_ptr<SetHandle> h = v[0]->handle();
        bool same = true; uint32_t n = kr_seqset_count(h->set);
        for (size_t i = 0; i < v.size(); ++i) same = same && v[i]->handle() == h && v[i]->index() == i;
        if (same && n == v.size()) return h;
        std::vector<uint32_t> idx; bool one = true;
        for (auto p : v) { one = one && p->handle() == h; idx.push_back(p->index()); }
        if (one) {
            kr_seqset *g = nullptr;
            KmerEngine::check(h->ctx, kr_seqset_gather(h->ctx, h->set, idx.data(), (uint32_t)idx.size(), &g));
            return std::make_shared<SetHandle>(h->ctx, g, h->K);
        }
        std::vector<const std::string *> s; for (auto p : v) s.push_back(&p->getSequence());
        return encodeStrings(v[0]->k(), s);
    }
    // bring the device copy of the database up to date: append new reps, refresh iteration ranks
    void sync() const {
        if (rebuild_ && devExists()) { devFree(); synced_ = 0; dev2host_.clear(); }
        if (!devExists()) devCreate();
        if (rebuild_) {   // replaced / removed reps: re-append the live ones
            FlatSeqs f;
            for (int i : order_.order()) { f.add(reps_[i]->getProtein()); dev2host_.push_back(i); }
            if (f.count()) devAppend(f);
            synced_ = reps_.size(); rebuild_ = false; ranks_dirty_ = true;
        } else if (synced_ < reps_.size()) {
            FlatSeqs f;
            for (size_t i = synced_; i < reps_.size(); ++i) { f.add(reps_[i]->getProtein()); dev2host_.push_back((int)i); }
            devAppend(f);
            synced_ = reps_.size(); ranks_dirty_ = true;
        }
        if (ranks_dirty_ && !dev2host_.empty()) {
            std::vector<int> ord = order_.order();
            std::vector<uint32_t> rank_of_host(reps_.size(), 0), rank(dev2host_.size());
            for (size_t k = 0; k < ord.size(); ++k) rank_of_host[ord[k]] = (uint32_t)k;
            for (size_t d = 0; d < dev2host_.size(); ++d) rank[d] = rank_of_host[dev2host_[d]];
            devSetRanks(rank);
        }
        ranks_dirty_ = false;
    }

    int threshold_, kmerSize_;
    std::string protName_;
    java::HashMapOrder order_;
    std::vector<std::unique_ptr<RepGenome>> reps_;   // by insertion index
    mutable kr_db *db_ = nullptr;
    mutable kr_multi_db *mdb_ = nullptr;
    mutable size_t synced_ = 0;                      // reps_[0..synced_) are on the device
    mutable std::vector<int> dev2host_;              // device rep index -> insertion index
    mutable bool rebuild_ = false, ranks_dirty_ = true;
    bool tie_last_ = false;
};

// ---------------------------------------------------------------------------------------------------
// org.theseed.proteins.kmers.KmerCollectionGroup (ClassifyProcessor.java:95,108,116): named groups of
// ProteinKmers; getDistance(seq, group) = smallest distance from seq to any member of the group
// (1.0 for an unknown or empty group).
class KmerCollectionGroup {
  public:
    void addSequence(const Sequence &seq, const std::string &group) { groups_[group].emplace_back(seq.sequence); dirty_.insert(group); }
    size_t size() const { return groups_.size(); }
    double getDistance(const Sequence &seq, const std::string &group) {
        std::vector<std::pair<const Sequence *, std::string>> one{{&seq, group}};
        return getDistances(one)[0];
    }
    // getBest(seq) [recalled member of the un-vendored class]: the group closest to the sequence and that distance; ties
    // go to the group that comes first in name order; ("", 1.0) when there is no group
    std::pair<std::string, double> getBest(const Sequence &seq) {
        std::vector<std::pair<const Sequence *, std::string>> q;
        for (auto &g : groups_) q.emplace_back(&seq, g.first);
        std::vector<double> d = getDistances(q);
        std::pair<std::string, double> best{"", 1.0};
        for (size_t i = 0; i < q.size(); ++i) if (d[i] < best.second) best = {q[i].second, d[i]};
        return best;
    }
    // batched form: every (sequence, group) query in one pass per group
    std::vector<double> getDistances(const std::vector<std::pair<const Sequence *, std::string>> &queries) {
        std::vector<double> out(queries.size(), 1.0);
        std::map<std::string, std::vector<size_t>> by_group;
        for (size_t i = 0; i < queries.size(); ++i) by_group[queries[i].second].push_back(i);
        for (auto &g : by_group) {
            auto it = groups_.find(g.first);
            if (it == groups_.end() || it->second.empty()) continue;
            auto &members = it->second;
            SequenceKmers::encodeBatch(members.begin(), members.end());
            std::vector<ProteinKmers> qs;
            for (size_t qi : g.second) qs.emplace_back(queries[qi].first->sequence);
            SequenceKmers::encodeBatch(qs.begin(), qs.end());
            std::vector<const std::string *> ms; for (auto &m : members) ms.push_back(&m.getProtein());
            auto hm = members[0].handle();
            bool same = true; for (size_t i = 0; i < members.size(); ++i) same = same && members[i].handle() == hm && members[i].index() == i;
            if (!same || kr_seqset_count(hm->set) != members.size()) hm = encodeStrings(members[0].k(), ms);
            auto hq = qs[0].handle();
            std::vector<int32_t> sim(qs.size() * members.size());
            kr_matrix_desc d{KR_ELEM_I32, KR_LAYOUT_FULL, 0, (uint32_t)qs.size()};
            KmerEngine::check(hq->ctx, kr_sim_matrix(hq->ctx, hq->set, hm->set, &d, sim.data()));
            for (size_t q = 0; q < qs.size(); ++q) {
                double best = 1.0;
                for (size_t m = 0; m < members.size(); ++m)
                    best = std::min(best, kr_distance(sim[q * members.size() + m], (uint32_t)qs[q].size(), (uint32_t)members[m].size()));
                out[g.second[q]] = best;
            }
        }
        return out;
    }
  private:
    std::map<std::string, std::vector<ProteinKmers>> groups_;
    std::set<std::string> dirty_;
};

}  // namespace theseed
