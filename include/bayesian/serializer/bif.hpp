// bayesian/serializer/bif.hpp — bn::serializer::bif: Bayesian Interchange Format -> (graph_t, database_t).
//
// Feeds the batched belief-propagation path with real network files (ALARM, HAILFINDER, ... as
// published in .bif form): parse -> bn::flatten (graph.hpp) -> bnbp_create.  Same class name, same
// `parse(begin, end)` signature and the same result as the reference loader
// (godai0519/BayesianNetwork, bayesian/serializer/bif.hpp:41-132), whose grammar is a Boost.Spirit.Qi
// grammar (:138-263).  Boost is not wanted here, so this is a hand-written recursive-descent parser
// of that grammar:
//
//   file        := ( network | variable | probability )*                               (:253-258)
//   network     := "network" name "{" "}"                                               (:220-223)
//   variable    := "variable" name "{" "type" "discrete" "[" uint "]"
//                  "{" name ("," name)* "}" ";" "}"                                     (:202-215)
//   probability := "probability" "(" name [ "|"+ name ("," name)* ] ")" "{" body "}"    (:228-248)
//   body        := [ ("table" numbers ";")+ | ( "(" name ("," name)* ")" numbers ";" )+ ]
//   numbers     := double ("," double)*
//
// What the reference does with the parse tree, kept (:56-131):
//   * one vertex per `variable` in file order, id = position, names into database_t;
//   * per `probability`: add_edge(parent, target) for the parents in the order listed, then
//     cpt.assign(parents, target) (every row zero) and one row per entry line, the state names of
//     the entry mapped to indices through the parents' own name lists;
//   * more than one `network` section is an error.
//
// Accepted beyond the reference grammar (real files need it; the reference throws on all of them):
// `//` and `/* */` comments, `property ... ;` lines inside network and variable blocks, quoted names,
// numbers separated by blanks instead of commas, `default` rows (fill every row not listed), and a
// `table` on a node WITH parents (BIF 0.15 order: the node's own state slowest, parents after it,
// last parent fastest).  Errors the reference leaves undefined (unknown variable or state name,
// a row of the wrong length) throw std::runtime_error here.
#ifndef BNB200_BAYESIAN_SERIALIZER_BIF_HPP
#define BNB200_BAYESIAN_SERIALIZER_BIF_HPP

#include <cstddef>
#include <fstream>
#include <iterator>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_map>
#include <utility>
#include <vector>

#include <bayesian/graph.hpp>
#include <bayesian/serializer/text_scanner.hpp>

namespace bn {
namespace serializer {

class bif {
public:
    // the parse tree, as in the reference (:24-38)
    struct variable_holder {
        std::string variable_name;
        unsigned int possible_value;
        std::vector<std::string> possible_value_name;
    };
    struct network_holder {
        std::string network_name;
    };
    struct probability_holder {
        typedef std::pair<std::string, std::vector<std::string>> rv_condition_type;
        typedef std::vector<std::pair<std::vector<std::string>, std::vector<double>>> probability_type;
        rv_condition_type rv_condition;
        probability_type probability;      // (state names of the parents, row); empty names = `table`
        std::vector<double> default_row;   // `default` line, if any
    };

    template <class Iterator>
    std::tuple<bn::graph_t, bn::database_t> parse(Iterator const& begin, Iterator const& end)
    {
        std::vector<network_holder> networks;
        std::vector<variable_holder> variables;
        std::vector<probability_holder> probabilities;
        text_scanner in(std::string(begin, end), "bif");
        while (!in.at_end()) {
            if (in.accept_word("network")) networks.push_back(network_section(in));
            else if (in.accept_word("variable")) variables.push_back(variable_section(in));
            else if (in.accept_word("probability")) probabilities.push_back(probability_section(in));
            else in.fail("expected 'network', 'variable' or 'probability'");
        }
        if (networks.size() > 1) throw std::runtime_error("too many network section in bif");

        bn::graph_t graph;
        bn::database_t database;
        if (networks.size() == 1) database.graph_name = networks[0].network_name;

        struct entry {
            bn::vertex_type vertex;
            std::vector<std::string> const* states;
        };
        std::unordered_map<std::string, entry> dictionary;
        for (variable_holder const& node : variables) {
            if (node.possible_value == 0 || node.possible_value != node.possible_value_name.size())
                throw std::runtime_error("cannot parse bif: variable " + node.variable_name +
                                         " declares a size that differs from its list of values");
            if (dictionary.count(node.variable_name))
                throw std::runtime_error("cannot parse bif: variable " + node.variable_name + " declared twice");
            bn::vertex_type vertex = graph.add_vertex();
            std::size_t const id = database.node_name.size();
            vertex->id = static_cast<int>(id);
            vertex->selectable_num = node.possible_value;
            database.node_name[id] = node.variable_name;
            database.options_name[id] = node.possible_value_name;
            dictionary[node.variable_name] = entry{vertex, &node.possible_value_name};
        }
        auto const lookup = [&dictionary](std::string const& name) -> entry const& {
            auto const at = dictionary.find(name);
            if (at == dictionary.end()) throw std::runtime_error("cannot parse bif: unknown variable " + name);
            return at->second;
        };
        auto const state_index = [](entry const& of, std::string const& value) -> int {
            for (std::size_t s = 0; s < of.states->size(); ++s)
                if ((*of.states)[s] == value) return static_cast<int>(s);
            throw std::runtime_error("cannot parse bif: unknown value " + value);
        };

        for (probability_holder const& probability : probabilities) {
            entry const& target = lookup(probability.rv_condition.first);
            std::vector<entry> parents;
            std::vector<bn::vertex_type> conditional_vertexes;
            for (std::string const& label : probability.rv_condition.second) {
                entry const& parent = lookup(label);
                if (!graph.add_edge(parent.vertex, target.vertex))
                    throw std::runtime_error("cannot parse bif: edge " + label + " -> " + probability.rv_condition.first +
                                             " repeats an edge or closes a cycle");
                parents.push_back(parent);
                conditional_vertexes.push_back(parent.vertex);
            }
            bn::cpt_t& cpt = target.vertex->cpt;
            cpt.assign(conditional_vertexes, target.vertex);
            std::size_t const width = target.vertex->selectable_num;
            auto const checked = [&](std::vector<double> const& row) -> std::vector<double> const& {
                if (row.size() != width)
                    throw std::runtime_error("cannot parse bif: a probability row of " + probability.rv_condition.first +
                                             " has " + std::to_string(row.size()) + " values, expected " + std::to_string(width));
                return row;
            };
            if (!probability.default_row.empty())
                for (bn::condition_t const& cond : cpt.pattern()) cpt[cond].second = checked(probability.default_row);
            for (auto const& line : probability.probability) {
                if (line.first.empty() && !parents.empty()) {          // `table` on a node with parents
                    fill_from_table(cpt, parents.size(), width, line.second, probability.rv_condition.first);
                    continue;
                }
                if (line.first.size() != parents.size())
                    throw std::runtime_error("cannot parse bif: a probability row of " + probability.rv_condition.first +
                                             " names " + std::to_string(line.first.size()) + " values for " +
                                             std::to_string(parents.size()) + " parents");
                bn::condition_t cond;
                for (std::size_t j = 0; j < parents.size(); ++j)
                    cond[parents[j].vertex] = state_index(parents[j], line.first[j]);
                cpt[cond].second = checked(line.second);
            }
        }
        return std::make_tuple(std::move(graph), std::move(database));
    }

    std::tuple<bn::graph_t, bn::database_t> from_data(std::string const& data) { return parse(data.begin(), data.end()); }

    std::tuple<bn::graph_t, bn::database_t> from_file(std::string const& filename)
    {
        std::ifstream ifs(filename, std::ios::binary);
        if (!ifs.is_open()) throw std::runtime_error("cannot open " + filename);
        std::string const data((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
        return parse(data.begin(), data.end());
    }

    // The inverse (new; the reference only reads): text this parser and the reference grammar both
    // accept, 17 significant digits so the CPTs survive the round trip bit for bit.
    static std::string write(bn::graph_t const& graph, bn::database_t const& database)
    {
        std::vector<bn::vertex_type> const& nodes = graph.vertex_list();
        auto const node_name = [&](std::size_t const i) {
            auto const at = database.node_name.find(i);
            return at != database.node_name.end() ? at->second : "n" + std::to_string(i);
        };
        auto const state_name = [&](std::size_t const i, std::size_t const s) {
            auto const at = database.options_name.find(i);
            return at != database.options_name.end() && s < at->second.size() ? at->second[s] : "s" + std::to_string(s);
        };
        std::ostringstream os;
        os.precision(17);
        os << "network " << (database.graph_name.empty() ? "unknown" : database.graph_name) << " {\n}\n";
        for (std::size_t i = 0; i < nodes.size(); ++i) {
            os << "variable " << node_name(i) << " {\n  type discrete [ " << nodes[i]->selectable_num << " ] { ";
            for (std::size_t s = 0; s < nodes[i]->selectable_num; ++s) os << (s ? ", " : "") << state_name(i, s);
            os << " };\n}\n";
        }
        for (std::size_t i = 0; i < nodes.size(); ++i) {
            std::vector<std::size_t> const parents = graph.parent_indices(i);
            os << "probability ( " << node_name(i);
            for (std::size_t j = 0; j < parents.size(); ++j) os << (j ? ", " : " | ") << node_name(parents[j]);
            os << " ) {\n";
            std::vector<std::size_t> state(parents.size(), 0);
            bn::cpt_t const& cpt = nodes[i]->cpt;
            for (bool more = true; more;) {
                bn::condition_t cond;
                for (std::size_t j = 0; j < parents.size(); ++j) cond[nodes[parents[j]]] = static_cast<int>(state[j]);
                auto const row = cpt[cond];
                if (!row.first) throw std::invalid_argument("bif::write: vertex " + std::to_string(i) + " lacks a CPT row");
                if (parents.empty()) os << "  table ";
                else {
                    os << "  (";
                    for (std::size_t j = 0; j < parents.size(); ++j) os << (j ? ", " : "") << state_name(parents[j], state[j]);
                    os << ") ";
                }
                for (std::size_t x = 0; x < row.second.size(); ++x) os << (x ? ", " : "") << row.second[x];
                os << ";\n";
                more = false;
                for (std::size_t j = parents.size(); j-- > 0;) {
                    if (++state[j] < nodes[parents[j]]->selectable_num) { more = true; break; }
                    state[j] = 0;
                }
            }
            os << "}\n";
        }
        return os.str();
    }

private:
    static void properties(text_scanner& in)
    {
        while (in.accept_word("property")) in.skip_past(';');
    }

    static network_holder network_section(text_scanner& in)
    {
        network_holder out;
        out.network_name = in.name();
        in.expect('{');
        properties(in);
        in.expect('}');
        return out;
    }

    static variable_holder variable_section(text_scanner& in)
    {
        variable_holder out;
        out.variable_name = in.name();
        in.expect('{');
        properties(in);
        in.expect_word("type");
        in.expect_word("discrete");
        in.expect('[');
        out.possible_value = static_cast<unsigned int>(in.unsigned_integer());
        in.expect(']');
        in.expect('{');
        while (in.peek() != '}') {                       // commas optional (JavaBayes omits them)
            out.possible_value_name.push_back(in.name());
            in.accept(',');
        }
        in.expect('}');
        in.expect(';');
        properties(in);
        in.expect('}');
        return out;
    }

    static std::vector<double> numbers(text_scanner& in)
    {
        std::vector<double> out;
        do {
            out.push_back(in.number());
            in.accept(',');
        } while (in.at_number());
        in.expect(';');
        return out;
    }

    static probability_holder probability_section(text_scanner& in)
    {
        probability_holder out;
        in.expect('(');
        out.rv_condition.first = in.name();
        if (in.accept('|')) {
            while (in.accept('|')) {}
            do out.rv_condition.second.push_back(in.name());
            while (in.accept(','));
        } else {
            // JavaBayes lists the parents after the node without a bar: probability ( "a" "b" "c" )
            while (in.peek() != ')') {
                in.accept(',');
                out.rv_condition.second.push_back(in.name());
            }
        }
        in.expect(')');
        in.expect('{');
        for (;;) {
            properties(in);
            if (in.accept('}')) break;
            if (in.accept_word("table")) {
                out.probability.push_back(std::make_pair(std::vector<std::string>(), numbers(in)));
            } else if (in.accept_word("default")) {
                out.default_row = numbers(in);
            } else {
                in.expect('(');
                std::vector<std::string> states;
                do states.push_back(in.name());
                while (in.accept(','));
                in.expect(')');
                out.probability.push_back(std::make_pair(states, numbers(in)));
            }
        }
        return out;
    }

    // BIF 0.15 `table`: all r_X * Q values, the node's own state slowest, then the parents in the
    // order listed, last parent fastest.
    static void fill_from_table(bn::cpt_t& cpt, std::size_t const n_parents, std::size_t const width,
                                std::vector<double> const& values, std::string const& who)
    {
        std::size_t const rows = cpt.row_count();
        if (values.size() != rows * width)
            throw std::runtime_error("cannot parse bif: table of " + who + " has " + std::to_string(values.size()) +
                                     " values, expected " + std::to_string(rows * width));
        // cpt.pattern() enumerates configurations in mixed radix over condition_node(), first parent
        // slowest: exactly the order of the table's parent digits
        std::vector<bn::condition_t> const keys = cpt.pattern();
        (void)n_parents;
        for (std::size_t q = 0; q < rows; ++q) {
            std::vector<double> row(width);
            for (std::size_t x = 0; x < width; ++x) row[x] = values[x * rows + q];
            cpt[keys[q]].second = row;
        }
    }
};

} // namespace serializer
} // namespace bn

#endif // BNB200_BAYESIAN_SERIALIZER_BIF_HPP
