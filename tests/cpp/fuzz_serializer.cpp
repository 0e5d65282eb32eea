// fuzz_serializer.cpp -- robustness of the network-file loaders (include/bayesian/serializer/{bif,dsc}.hpp, the front of
// SURVEY section 8 f1) under AddressSanitizer + UndefinedBehaviorSanitizer: every truncation of a valid file and a few
// thousand single-byte mutations must either load or throw std::exception -- never read out of bounds, overflow or abort.
// Host code only; built and run by tests/test_sanitizers.py (-m "not gpu").
//
//   fuzz_serializer <file.bif|file.dsc> [mutations]
#include <cstdint>
#include <cstdio>
#include <exception>
#include <fstream>
#include <iostream>
#include <iterator>
#include <string>

#include "bayesian/graph.hpp"
#include "bayesian/serializer/bif.hpp"
#include "bayesian/serializer/dsc.hpp"

namespace {

struct tally { long loaded = 0, refused = 0; };

void feed(std::string const& text, bool dsc, tally& t)
{
    try {
        if (dsc) {
            bn::graph_t g = bn::serializer::dsc().from_data(text);
            (void)bn::flatten(g);
        } else {
            auto r = bn::serializer::bif().parse(text.begin(), text.end());
            (void)bn::flatten(std::get<0>(r));
        }
        ++t.loaded;
    } catch (std::exception const&) {
        ++t.refused;
    }
}

std::uint64_t next(std::uint64_t& s)
{
    s += 0x9E3779B97F4A7C15ull;
    std::uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

} // namespace

int main(int argc, char** argv)
{
    if (argc < 2) { std::fprintf(stderr, "usage: fuzz_serializer <file> [mutations]\n"); return 2; }
    std::string const path = argv[1];
    bool const dsc = path.size() > 4 && path.substr(path.size() - 4) == ".dsc";
    long const mutations = argc > 2 ? std::atol(argv[2]) : 2000;
    std::ifstream in(path, std::ios::binary);
    std::string const text((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (text.empty()) { std::fprintf(stderr, "cannot read %s\n", path.c_str()); return 2; }

    tally whole, cut, mut;
    feed(text, dsc, whole);
    if (whole.loaded != 1) { std::fprintf(stderr, "the unmodified file does not load\n"); return 1; }
    // every prefix (stride keeps the run in seconds on the larger files)
    std::size_t const stride = text.size() > 4000 ? text.size() / 4000 : 1;
    for (std::size_t n = 0; n < text.size(); n += stride) feed(text.substr(0, n), dsc, cut);
    // single-byte mutations: structural characters, digits, NUL and high bytes at random places
    static char const pool[] = "{}()[]|,;:\"=. \n0919-+eE*/#\0\xff";
    std::uint64_t seed = 20261018;
    for (long i = 0; i < mutations; ++i) {
        std::string m = text;
        std::size_t const at = (std::size_t)(next(seed) % m.size());
        switch (next(seed) % 3) {
            case 0: m[at] = pool[next(seed) % (sizeof pool - 1)]; break;
            case 1: m.erase(at, 1 + (std::size_t)(next(seed) % 7)); break;
            default: m.insert(at, 1, pool[next(seed) % (sizeof pool - 1)]); break;
        }
        feed(m, dsc, mut);
    }
    std::cout << "prefixes: " << cut.loaded << " loaded, " << cut.refused << " refused; mutations: " << mut.loaded
              << " loaded, " << mut.refused << " refused\n";
    return 0;
}
