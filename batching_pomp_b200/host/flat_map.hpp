// flat_map.hpp — open-addressing hash table for 64-bit keys (linear probing, power-of-two capacity,
// tombstones).  The host planner keeps ~10^7 roadmap edges indexed by vertex pair; a node-based
// std::unordered_map costs an allocation and a cache miss per insert, this costs one cache miss.
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

namespace batching_pomp {
namespace util {

template <class V>
class FlatMap64 {
public:
  FlatMap64() : mSize(0), mUsed(0) { rehash(16); }

  size_t size() const { return mSize; }

  void clear() {
    std::fill(mState.begin(), mState.end(), (uint8_t)0);
    mSize = mUsed = 0;
  }

  void reserve(size_t n) {
    size_t want = 16;
    while (want * 7 < n * 10) want <<= 1;  // load factor <= 0.7
    if (want > mState.size()) rehash(want);
  }

  /// Pointer to the value of `key`, or nullptr.
  V* find(uint64_t key) {
    size_t mask = mState.size() - 1, i = hash(key) & mask;
    while (mState[i] != 0) {
      if (mState[i] == 1 && mKeys[i] == key) return &mVals[i];
      i = (i + 1) & mask;
    }
    return nullptr;
  }
  const V* find(uint64_t key) const { return const_cast<FlatMap64*>(this)->find(key); }

  /// Inserts (key, value) unless the key is present (std::unordered_map::emplace semantics).
  /// Returns true if inserted.
  bool emplace(uint64_t key, const V& value) {
    if ((mUsed + 1) * 10 > mState.size() * 7) rehash(mSize * 10 > mState.size() * 3 ? mState.size() * 2 : mState.size());
    size_t mask = mState.size() - 1, i = hash(key) & mask, tomb = (size_t)-1;
    while (mState[i] != 0) {
      if (mState[i] == 1 && mKeys[i] == key) return false;
      if (mState[i] == 2 && tomb == (size_t)-1) tomb = i;
      i = (i + 1) & mask;
    }
    if (tomb != (size_t)-1) i = tomb;
    else ++mUsed;
    mState[i] = 1;
    mKeys[i] = key;
    mVals[i] = value;
    ++mSize;
    return true;
  }

  bool erase(uint64_t key) {
    size_t mask = mState.size() - 1, i = hash(key) & mask;
    while (mState[i] != 0) {
      if (mState[i] == 1 && mKeys[i] == key) {
        mState[i] = 2;
        --mSize;
        return true;
      }
      i = (i + 1) & mask;
    }
    return false;
  }

private:
  static size_t hash(uint64_t k) {  // splitmix64 finaliser
    k ^= k >> 30;
    k *= 0xbf58476d1ce4e5b9ull;
    k ^= k >> 27;
    k *= 0x94d049bb133111ebull;
    k ^= k >> 31;
    return (size_t)k;
  }
  void rehash(size_t cap) {
    std::vector<uint8_t> st(cap, 0);
    std::vector<uint64_t> ks(cap);
    std::vector<V> vs(cap);
    size_t mask = cap - 1;
    for (size_t j = 0; j < mState.size(); ++j)
      if (mState[j] == 1) {
        size_t i = hash(mKeys[j]) & mask;
        while (st[i] != 0) i = (i + 1) & mask;
        st[i] = 1;
        ks[i] = mKeys[j];
        vs[i] = mVals[j];
      }
    mState.swap(st);
    mKeys.swap(ks);
    mVals.swap(vs);
    mUsed = mSize;
  }

  std::vector<uint8_t> mState;  // 0 empty, 1 full, 2 tombstone
  std::vector<uint64_t> mKeys;
  std::vector<V> mVals;
  size_t mSize, mUsed;
};

}  // namespace util
}  // namespace batching_pomp
