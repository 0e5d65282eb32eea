// Stand-in for <boost/algorithm/string.hpp> (Boost is not in this image) — TEST INFRASTRUCTURE.
// Exactly what bayesian/sampler.hpp:53-54 of the reference uses: boost::algorithm::split(result, input,
// boost::is_space(), boost::algorithm::token_compress_on), plus trim / is_any_of for completeness.  Semantics of
// the real thing: the input is cut at every character the predicate accepts; with token_compress_on adjacent
// delimiters count as one; leading / trailing delimiters yield EMPTY first / last tokens (Boost does that too —
// the reference's sample files start every line with the count, so line[0] is the count).
#ifndef BNB200_TESTS_BOOST_ALGORITHM_STRING_HPP
#define BNB200_TESTS_BOOST_ALGORITHM_STRING_HPP
#include <cctype>
#include <string>
namespace boost {
namespace algorithm {
enum token_compress_mode_type { token_compress_on, token_compress_off };
struct is_space_pred { bool operator()(char ch) const { return std::isspace(static_cast<unsigned char>(ch)) != 0; } };
struct is_any_of_pred {
    std::string set;
    bool operator()(char ch) const { return set.find(ch) != std::string::npos; }
};
inline is_space_pred is_space() { return is_space_pred(); }
inline is_any_of_pred is_any_of(std::string const& s) { is_any_of_pred p; p.set = s; return p; }

template <class Container, class Pred>
Container& split(Container& result, std::string const& input, Pred pred, token_compress_mode_type mode = token_compress_off)
{
    result.clear();
    std::string token;
    bool last_was_delim = false;
    for (std::string::size_type i = 0; i < input.size(); ++i) {
        if (pred(input[i])) {
            if (!(mode == token_compress_on && last_was_delim)) { result.push_back(token); token.clear(); }
            last_was_delim = true;
        } else {
            token.push_back(input[i]);
            last_was_delim = false;
        }
    }
    result.push_back(token);
    return result;
}

inline void trim(std::string& s)
{
    std::string::size_type a = 0, b = s.size();
    while (a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while (b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    s = s.substr(a, b - a);
}
inline std::string trim_copy(std::string s) { trim(s); return s; }
} // namespace algorithm
using algorithm::is_any_of;
using algorithm::is_space;
using algorithm::split;
using algorithm::trim;
} // namespace boost
#endif
