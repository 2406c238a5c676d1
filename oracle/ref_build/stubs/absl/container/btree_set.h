// ORACLE SCAFFOLDING: std::set standing in for absl::btree_set (same ordered-set semantics).
#pragma once
#include <set>
namespace absl {
template <class K, class C = std::less<K>>
class btree_set : public std::set<K, C> {
    using B = std::set<K, C>;
   public:
    using B::B;
    bool contains(const K& k) const { return this->find(k) != this->end(); }
};
}  // namespace absl
