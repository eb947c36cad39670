// Host side of the two programs: text formats, libstdc++ hash-order emulation, per-call annotations
// and the argv-compatible mains.  (Round 1: hash order only; loaders/writers/mains follow.)
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>

#include "../../include/amplisolve_b200.h"

extern "C" {

// The reference walks std::unordered_map<std::string,std::string> containers (EE:1081 normals, VC:672
// tumours, VC:1046 FILTER flags) and that order is visible in its outputs.  The only faithful model of
// libstdc++'s order is libstdc++: insert the same keys in the same sequence, read the order back.
int as_hash_iteration_order(const char* const* keys, int32_t n, int32_t* order_out) {
    if (!keys || !order_out || n < 0) return AS_EINVAL;
    std::unordered_map<std::string, std::string> m;
    std::unordered_map<std::string, int32_t> first;
    for (int32_t i = 0; i < n; ++i) {
        m.insert(std::make_pair(std::string(keys[i]), std::string()));
        first.insert(std::make_pair(std::string(keys[i]), i));
    }
    int32_t j = 0;
    for (auto it = m.begin(); it != m.end(); ++it) order_out[j++] = first[it->first];
    return j;
}

}  // extern "C"
