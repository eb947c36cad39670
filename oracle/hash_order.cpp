// TEST INFRASTRUCTURE ONLY -- part of liboracle.so.
// The reference iterates std::unordered_map<std::string,std::string> containers whose
// iteration order is visible in its outputs: normal-file order (EE:1081, matters for Germ_Max),
// tumour-file order (VC:672) and FILTER-flag order (VC:1046-1059).  The only faithful model of
// that order is libstdc++ itself, so the oracle asks it: insert the same keys in the same
// sequence, read the order back.
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>

extern "C" void aso_hash_iteration_order(const char* const* keys, int n, int32_t* order_out) {
    std::unordered_map<std::string, std::string> m;
    std::unordered_map<std::string, int> first_index;
    for (int i = 0; i < n; ++i) {
        m.insert(std::make_pair(std::string(keys[i]), std::string(keys[i])));  // insert: first wins
        first_index.insert(std::make_pair(std::string(keys[i]), i));
    }
    int j = 0;
    for (std::unordered_map<std::string, std::string>::iterator it = m.begin(); it != m.end(); ++it)
        order_out[j++] = first_index[it->first];
    for (; j < n; ++j) order_out[j] = -1;
}
