// ORACLE SCAFFOLDING (test infrastructure, not product code).
// The reference links TurakhiaLab/panman v0.1.4 (CMakeLists.txt:317-322), which is fetched at
// configure time and is absent from this image.  This header declares just the data model the
// reference's own index builder touches (Tree / Node / NucMut / BlockMut / Block / GapList) so that
// /root/reference/src/index_single_mode.cpp and panmap_utils.cpp compile UNMODIFIED against it.
// oracle/ref_build/panman_loader.hpp fills these structs from a .panman file.
#pragma once
#include <cstdint>
#include <istream>
#include <string>
#include <unordered_map>
#include <vector>
namespace panmanUtils {
enum NucMutationType { NS = 0, ND = 1, NI = 2, NSNPS = 3, NSNPI = 4, NSNPD = 5 };
struct NucMut {
    int32_t nucPosition = 0;
    int32_t nucGapPosition = -1;
    int32_t primaryBlockId = 0;
    int32_t secondaryBlockId = -1;
    uint8_t mutInfo = 0;   // (length << 4) | type
    uint32_t nucs = 0;     // up to 6 four-bit codes, first base in bits 20..23
};
struct BlockMut {
    int32_t primaryBlockId = 0;
    int32_t secondaryBlockId = -1;
    bool blockMutInfo = false;  // true = insertion
    bool inversion = false;
};
struct Block {
    int64_t primaryBlockId = 0;
    int64_t secondaryBlockId = -1;
    std::vector<uint32_t> consensusSeq;
    std::string chromosomeName;
};
struct GapList {
    std::vector<int32_t> nucPosition;
    std::vector<int32_t> nucGapLength;
    int32_t primaryBlockId = 0;
    int32_t secondaryBlockId = -1;
};
class Node {
   public:
    float branchLength = 0.f;
    size_t level = 0;
    std::string identifier;
    Node* parent = nullptr;
    std::vector<Node*> children;
    std::vector<NucMut> nucMutation;
    std::vector<BlockMut> blockMutation;
    std::vector<std::string> annotations;
};
class Tree {
   public:
    Node* root = nullptr;
    std::unordered_map<std::string, Node*> allNodes;
    std::vector<Block> blocks;
    std::vector<GapList> gaps;
    ~Tree() { for (auto& kv : allNodes) delete kv.second; }
    Tree() = default;
    Tree(const Tree&) = delete;
    Tree& operator=(const Tree&) = delete;
    // only reached by the alignment-refinement path (off by default, out of scope for the oracle)
    std::string getStringFromReference(const std::string&, bool, bool) { return {}; }
};
inline char getNucleotideFromCode(int code) {
    switch (code) {
        case 1: return 'A'; case 2: return 'C'; case 4: return 'G'; case 8: return 'T';
        case 5: return 'R'; case 10: return 'Y'; case 6: return 'S'; case 9: return 'W';
        case 12: return 'K'; case 3: return 'M'; case 14: return 'B'; case 13: return 'D';
        case 11: return 'H'; case 7: return 'V'; case 15: return 'N';
        default: return '-';
    }
}
inline char getComplementCharacter(char c) {
    switch (c) {
        case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
        case 'R': return 'Y'; case 'Y': return 'R'; case 'K': return 'M'; case 'M': return 'K';
        case 'B': return 'V'; case 'V': return 'B'; case 'D': return 'H'; case 'H': return 'D';
        default: return c;  // S, W, N, '-'
    }
}
}  // namespace panmanUtils
