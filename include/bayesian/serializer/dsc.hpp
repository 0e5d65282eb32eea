// bayesian/serializer/dsc.hpp — bn::serializer::dsc: Microsoft Research DSC network files -> graph_t.
//
// The second text format the reference reads (godai0519/BayesianNetwork,
// bayesian/serializer/dsc.hpp:33-232).  Same class, same entry points (`parse(lines)`,
// `from_file`, `from_data`, free function `stream_to_lines`) and the same resulting graph:
//   * `node NAME { ... type: discrete[n] = {...}; ... }`  -> add_vertex, selectable_num = n (:116-150);
//     every other attribute of the block is ignored, as in the reference;
//   * `probability(X | A, B) { (i, j): p, ...; ... }`     -> add_edge(A, X), add_edge(B, X) in the
//     order listed, cpt.assign({A, B}, X), one row per line, parent states given as indices (:152-226);
//   * `probability(X) { p, ...; }`                        -> the single prior row (:180-192);
//   * `belief network "name"`                             -> header; the reference drops the name
//     (:104-114), here it is kept in database().graph_name.
// The reference reads whole lines at fixed column offsets (`it->substr(15)`, `substr(12)`: a second
// blank or a brace on the same line breaks it); this is a token-level parser of the same format, so
// layout does not matter.  Also accepted: `default: ...;` rows, state NAMES instead of indices in a
// condition tuple, comments, attributes spread over several lines.  Unknown variables and rows of
// the wrong length throw std::runtime_error (undefined behaviour in the reference).
#ifndef BNB200_BAYESIAN_SERIALIZER_DSC_HPP
#define BNB200_BAYESIAN_SERIALIZER_DSC_HPP

#include <cctype>
#include <cstddef>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include <bayesian/graph.hpp>
#include <bayesian/serializer/text_scanner.hpp>

namespace bn {
namespace serializer {

template <class InputStream>
std::vector<std::string> stream_to_lines(InputStream& is)
{
    std::vector<std::string> result;
    std::string line;
    while (std::getline(is, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        result.push_back(line);
    }
    return result;
}

class dsc {
public:
    graph_t parse(std::vector<std::string> const& data)
    {
        std::string text;
        for (std::string const& line : data) {
            text += line;
            text += '\n';
        }
        return parse_text(text);
    }

    graph_t from_file(std::string const& filename)
    {
        std::ifstream ifs(filename);
        if (!ifs.is_open()) throw std::runtime_error("cannot open " + filename);
        return parse(stream_to_lines(ifs));
    }

    graph_t from_data(std::string const& data)
    {
        std::istringstream iss(data);
        return parse(stream_to_lines(iss));
    }

    // names of the last parse (new: the reference forgets them)
    database_t const& database() const { return database_; }

private:
    struct entry {
        vertex_type vertex;
        std::size_t index;
    };

    graph_t parse_text(std::string const& text)
    {
        graph_t graph;
        database_ = database_t();
        dictionary_.clear();
        text_scanner in(text, "dsc");
        while (!in.at_end()) {
            if (in.accept_word("belief")) {
                in.expect_word("network");
                database_.graph_name = in.name();
            } else if (in.accept_word("node")) {
                node_section(in, graph);
            } else if (in.accept_word("probability")) {
                probability_section(in, graph);
            } else if (in.accept_word("properties") || in.accept_word("domain")) {
                skip_block(in);
            } else {
                in.fail("expected 'belief network', 'node' or 'probability'");
            }
        }
        return graph;
    }

    static void skip_block(text_scanner& in)
    {
        while (in.peek() != '{') in.name();
        in.expect('{');
        int depth = 1;
        while (depth > 0) {
            if (in.at_end()) in.fail("unterminated block");
            if (in.accept('{')) ++depth;
            else if (in.accept('}')) --depth;
            else if (in.peek() == '"') in.name();
            else in.skip_past(in.peek());
        }
    }

    void node_section(text_scanner& in, graph_t& graph)
    {
        std::string const node_name = in.name();
        if (dictionary_.count(node_name)) in.fail("node " + node_name + " declared twice");
        vertex_type vertex = graph.add_vertex();
        std::size_t const index = dictionary_.size();
        vertex->id = static_cast<int>(index);
        dictionary_[node_name] = entry{vertex, index};
        database_.node_name[index] = node_name;
        in.expect('{');
        while (!in.accept('}')) {
            std::string const key = in.name();
            if (!in.accept(':')) in.accept('=');
            if (key == "type") {
                in.expect_word("discrete");
                in.expect('[');
                vertex->selectable_num = in.unsigned_integer();
                in.expect(']');
                std::vector<std::string> states;
                if (in.accept('=')) {
                    in.expect('{');
                    while (!in.accept('}')) {
                        states.push_back(in.name());
                        in.accept(',');
                    }
                }
                in.expect(';');
                if (!states.empty() && states.size() != vertex->selectable_num)
                    in.fail("node " + node_name + " declares a size that differs from its list of values");
                database_.options_name[index] = states;
            } else {
                skip_attribute(in);
            }
        }
    }

    // `name: "..."; position: (1, 2); ...` — everything up to the terminating ';' outside quotes
    static void skip_attribute(text_scanner& in)
    {
        for (;;) {
            char const c = in.peek();
            if (c == '\0') in.fail("unterminated attribute");
            if (c == '"') { in.name(); continue; }
            in.skip_past(c);
            if (c == ';') return;
        }
    }

    void probability_section(text_scanner& in, graph_t& graph)
    {
        in.expect('(');
        std::string const node_name = in.name();
        entry const target = lookup(in, node_name);
        std::vector<entry> parents;
        std::vector<vertex_type> cond_nodes;
        if (in.accept('|')) {
            do {
                std::string const parent_name = in.name();
                entry const parent = lookup(in, parent_name);
                if (!graph.add_edge(parent.vertex, target.vertex))
                    in.fail("edge " + parent_name + " -> " + node_name + " repeats an edge or closes a cycle");
                parents.push_back(parent);
                cond_nodes.push_back(parent.vertex);
            } while (in.accept(','));
        }
        in.expect(')');
        cpt_t& cpt = target.vertex->cpt;
        cpt.assign(cond_nodes, target.vertex);
        std::size_t const width = target.vertex->selectable_num;
        in.expect('{');
        while (!in.accept('}')) {
            if (in.accept_word("default")) {
                in.expect(':');
                std::vector<double> const row = numbers(in, width, node_name);
                for (condition_t const& cond : cpt.pattern()) cpt[cond].second = row;
            } else if (in.accept('(')) {
                condition_t cond;
                for (std::size_t j = 0; j < parents.size(); ++j) {
                    if (j) in.expect(',');
                    cond[parents[j].vertex] = state_of(in, parents[j]);
                }
                in.expect(')');
                in.expect(':');
                auto slot = cpt[cond];
                if (!slot.first) in.fail("a condition of " + node_name + " is out of range");
                slot.second = numbers(in, width, node_name);
            } else {
                if (!parents.empty()) in.fail("a row of " + node_name + " lacks its condition tuple");
                cpt[condition_t()].second = numbers(in, width, node_name);
            }
        }
    }

    int state_of(text_scanner& in, entry const& parent) const
    {
        std::string const token = in.name();
        bool digits = !token.empty();
        for (char const c : token) digits = digits && std::isdigit(static_cast<unsigned char>(c));
        if (digits) return std::stoi(token);            // the reference's form: state indices (:217-220)
        auto const names = database_.options_name.find(parent.index);
        if (names != database_.options_name.end())
            for (std::size_t s = 0; s < names->second.size(); ++s)
                if (names->second[s] == token) return static_cast<int>(s);
        in.fail("unknown value " + token);
    }

    static std::vector<double> numbers(text_scanner& in, std::size_t const width, std::string const& who)
    {
        std::vector<double> out;
        do {
            out.push_back(in.number());
            in.accept(',');
        } while (in.at_number());
        in.expect(';');
        if (out.size() != width)
            in.fail("a probability row of " + who + " has " + std::to_string(out.size()) + " values, expected " +
                    std::to_string(width));
        return out;
    }

    entry lookup(text_scanner& in, std::string const& name) const
    {
        auto const at = dictionary_.find(name);
        if (at == dictionary_.end()) in.fail("unknown node " + name);
        return at->second;
    }

    database_t database_;
    std::unordered_map<std::string, entry> dictionary_;
};

} // namespace serializer
} // namespace bn

#endif // BNB200_BAYESIAN_SERIALIZER_DSC_HPP
