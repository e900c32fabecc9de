// Test infrastructure: C entry point over the reference's own paths-file parser (hevc/hm_common/c++/source_common/tools.cpp,
// compiled unmodified next to this file by oracle/Makefile) and the selection TComPrediction::initTempBuff makes from its
// two maps (TComPrediction.cpp(substitution):145-171).  Returns 1 and the chosen path, 0 when the key is absent, -1 when
// the parser fails or throws.
#include <cstring>
#include <map>
#include <string>

#include "tools.h"

extern "C" int ref_choose_path(const char* path_to_file, int qp_selection, unsigned int width, char* out, int out_size) {
    try {
        std::map<std::pair<unsigned int, unsigned int>, std::string> single, pair;
        if (parse_file_strings_three_keys(single, pair, path_to_file, ",") < 0) return -1;
        const std::map<std::pair<unsigned int, unsigned int>, std::string>& chosen = (!pair.empty() && qp_selection >= 32) ? pair : single;
        const auto it = chosen.find(std::make_pair(width, 0u));
        if (it == chosen.end()) return 0;
        if ((int)it->second.size() + 1 > out_size) return -1;
        std::memcpy(out, it->second.c_str(), it->second.size() + 1);
        return 1;
    } catch (...) {
        return -1;
    }
}
