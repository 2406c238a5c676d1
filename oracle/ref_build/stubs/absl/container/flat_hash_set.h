// ORACLE SCAFFOLDING: std::unordered_set standing in for absl::flat_hash_set.
#pragma once
#include <functional>
#include <unordered_set>
namespace absl {
template <class K, class H = std::hash<K>, class E = std::equal_to<K>>
class flat_hash_set : public std::unordered_set<K, H, E> {
    using B = std::unordered_set<K, H, E>;
   public:
    using B::B;
    bool contains(const K& k) const { return this->find(k) != this->end(); }
};
}  // namespace absl
