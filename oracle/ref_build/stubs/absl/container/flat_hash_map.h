// ORACLE SCAFFOLDING: std::unordered_map standing in for absl::flat_hash_map (abseil is absent here).
// Only iteration order differs (affects f64 summation order of the read-magnitude sums, nothing integer).
#pragma once
#include <functional>
#include <unordered_map>
namespace absl {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>>
class flat_hash_map : public std::unordered_map<K, V, H, E> {
    using B = std::unordered_map<K, V, H, E>;
   public:
    using B::B;
    void prefetch(const K&) const {}
    bool contains(const K& k) const { return this->find(k) != this->end(); }
};
}  // namespace absl
