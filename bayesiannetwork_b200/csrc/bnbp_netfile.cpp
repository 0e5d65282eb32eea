// bnbp_netfile.cpp — network files behind the C ABI: BIF / DSC text ->
// bnbp_flat_network.  The parsers are the drop-in C++ headers (include/bayesian/serializer/*.hpp,
// replacing the reference's bayesian/serializer/bif.hpp:41-132 and dsc.hpp:33-232); this file only
// flattens their graph_t (bn::flatten, include/bayesian/graph.hpp) and hands out plain arrays, so that
// FFI hosts (the Python mirror, bench.py) load the same files through the same code.  Host-only.
#include <bnbp.h>

#include <bayesian/graph.hpp>
#include <bayesian/serializer/bif.hpp>
#include <bayesian/serializer/dsc.hpp>

#include <cctype>
#include <cstring>
#include <fstream>
#include <iterator>
#include <memory>
#include <string>
#include <vector>

namespace bnbp {
int set_error(int code, const std::string& msg);   // bnbp_api.cu
}

struct bnbp_network_file {
    bn::flat_network flat;
    bnbp_flat_network view;
    std::string name;
    std::vector<std::string> node_name;
    std::vector<std::vector<std::string>> state_name;
};

namespace {

bool ends_with(const std::string& s, const char* tail)
{
    const size_t n = strlen(tail);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i)
        if (tolower((unsigned char)s[s.size() - n + i]) != tail[i]) return false;
    return true;
}

int sniff(const std::string& text)
{
    // DSC files declare `node X`, BIF files `variable X`
    const size_t v = text.find("variable"), n = text.find("node");
    if (v != std::string::npos && (n == std::string::npos || v < n)) return BNBP_FORMAT_BIF;
    if (n != std::string::npos) return BNBP_FORMAT_DSC;
    return BNBP_FORMAT_BIF;
}

int build(const std::string& text, int format, bnbp_network_file** out)
{
    if (!out) return bnbp::set_error(BNBP_ERR_INVALID, "bnbp_netfile: NULL output pointer");
    *out = nullptr;
    if (format == BNBP_FORMAT_AUTO) format = sniff(text);
    try {
        bn::graph_t graph;
        bn::database_t names;
        if (format == BNBP_FORMAT_BIF) {
            bn::serializer::bif reader;
            std::tie(graph, names) = reader.parse(text.begin(), text.end());
        } else if (format == BNBP_FORMAT_DSC) {
            bn::serializer::dsc reader;
            graph = reader.from_data(text);
            names = reader.database();
        } else {
            return bnbp::set_error(BNBP_ERR_INVALID, "bnbp_netfile: unknown format");
        }
        std::unique_ptr<bnbp_network_file> nf(new bnbp_network_file());
        nf->flat = bn::flatten(graph);
        nf->name = names.graph_name;
        const size_t n = nf->flat.card.size();
        nf->node_name.resize(n);
        nf->state_name.resize(n);
        for (size_t i = 0; i < n; ++i) {
            auto nm = names.node_name.find(i);
            nf->node_name[i] = nm != names.node_name.end() ? nm->second : "n" + std::to_string(i);
            auto st = names.options_name.find(i);
            if (st != names.options_name.end()) nf->state_name[i] = st->second;
            for (size_t s = nf->state_name[i].size(); s < (size_t)nf->flat.card[i]; ++s)
                nf->state_name[i].push_back(std::to_string(s));
        }
        nf->view.n_nodes = (int32_t)n;
        nf->view.card = nf->flat.card.data();
        nf->view.parent_off = nf->flat.parent_off.data();
        nf->view.parents = nf->flat.parents.data();
        nf->view.cpt_off = nf->flat.cpt_off.data();
        nf->view.cpt = nf->flat.cpt.data();
        *out = nf.release();
        return BNBP_OK;
    } catch (const std::exception& e) {
        return bnbp::set_error(BNBP_ERR_INVALID, e.what());
    }
}

}  // namespace

extern "C" {

int bnbp_netfile_parse(const char* text, int64_t len, int32_t format, bnbp_network_file** out)
{
    if (!text || len < 0) return bnbp::set_error(BNBP_ERR_INVALID, "bnbp_netfile_parse: NULL text");
    return build(std::string(text, (size_t)len), format, out);
}

int bnbp_netfile_load(const char* path, int32_t format, bnbp_network_file** out)
{
    if (!path) return bnbp::set_error(BNBP_ERR_INVALID, "bnbp_netfile_load: NULL path");
    std::ifstream ifs(path, std::ios::binary);
    if (!ifs.is_open()) return bnbp::set_error(BNBP_ERR_INVALID, std::string("cannot open ") + path);
    const std::string text((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
    if (format == BNBP_FORMAT_AUTO) {
        if (ends_with(path, ".bif")) format = BNBP_FORMAT_BIF;
        else if (ends_with(path, ".dsc")) format = BNBP_FORMAT_DSC;
    }
    return build(text, format, out);
}

const bnbp_flat_network* bnbp_netfile_network(const bnbp_network_file* nf) { return nf ? &nf->view : nullptr; }

const char* bnbp_netfile_name(const bnbp_network_file* nf) { return nf ? nf->name.c_str() : ""; }

const char* bnbp_netfile_node_name(const bnbp_network_file* nf, int32_t node)
{
    if (!nf || node < 0 || (size_t)node >= nf->node_name.size()) return nullptr;
    return nf->node_name[(size_t)node].c_str();
}

const char* bnbp_netfile_state_name(const bnbp_network_file* nf, int32_t node, int32_t state)
{
    if (!nf || node < 0 || (size_t)node >= nf->state_name.size()) return nullptr;
    const std::vector<std::string>& names = nf->state_name[(size_t)node];
    if (state < 0 || (size_t)state >= names.size()) return nullptr;
    return names[(size_t)state].c_str();
}

void bnbp_netfile_free(bnbp_network_file* nf) { delete nf; }

}  // extern "C"
