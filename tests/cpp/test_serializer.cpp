// Own tests of the drop-in loaders include/bayesian/serializer/{bif,dsc}.hpp:
// the classes are used exactly as the reference's are (bn::serializer::bif().parse(begin, end) ->
// tuple<graph_t, database_t>; bn::serializer::dsc().from_data(text) -> graph_t).  Host code only.
#define BOOST_TEST_MAIN
#include <boost/test/included/unit_test.hpp>

#include <string>
#include <tuple>

#include "bayesian/graph.hpp"
#include "bayesian/serializer/bif.hpp"
#include "bayesian/serializer/dsc.hpp"

namespace {

char const* const kBif =
    "network sprinkler {\n}\n"
    "variable Cloudy {\n  type discrete [ 2 ] { no, yes };\n}\n"
    "variable Rain {\n  type discrete [ 3 ] { none, light, heavy };\n}\n"
    "variable Wet {\n  type discrete [ 2 ] { dry, wet };\n}\n"
    "probability ( Cloudy ) {\n  table 0.5, 0.5;\n}\n"
    "probability ( Rain | Cloudy ) {\n  (no) 0.8, 0.15, 0.05;\n  (yes) 0.2, 0.5, 0.3;\n}\n"
    "probability ( Wet | Rain, Cloudy ) {\n"
    "  (none, no) 1.0, 0.0;\n  (none, yes) 0.9, 0.1;\n  (light, no) 0.3, 0.7;\n"
    "  (light, yes) 0.2, 0.8;\n  (heavy, no) 0.1, 0.9;\n  (heavy, yes) 0.05, 0.95;\n}\n";

char const* const kDsc =
    "belief network \"sprinkler\"\n"
    "node Cloudy\n{\n  name: \"Cloudy\";\n  type: discrete[2] = {\"no\", \"yes\"};\n}\n"
    "node Rain\n{\n  type: discrete[3] = {\"none\", \"light\", \"heavy\"};\n}\n"
    "probability(Cloudy)\n{\n  0.5, 0.5;\n}\n"
    "probability(Rain | Cloudy)\n{\n  (0): 0.8, 0.15, 0.05;\n  (1): 0.2, 0.5, 0.3;\n}\n";

} // namespace

BOOST_AUTO_TEST_CASE( bif_parse_like_the_reference )
{
    std::string const text(kBif);
    bn::graph_t graph;
    bn::database_t db;
    std::tie(graph, db) = bn::serializer::bif().parse(text.begin(), text.end());
    auto const& v = graph.vertex_list();
    BOOST_CHECK(v.size() == 3);
    BOOST_CHECK(db.graph_name == "sprinkler");
    BOOST_CHECK(db.node_name[1] == "Rain");
    BOOST_CHECK(db.options_name[1].size() == 3 && db.options_name[1][2] == "heavy");
    BOOST_CHECK(v[0]->id == 0 && v[2]->id == 2);
    BOOST_CHECK(v[1]->selectable_num == 3);
    BOOST_CHECK(graph.edge_list().size() == 3);
    auto const parents = graph.in_vertexes(v[2]);          // ascending index whatever the file's order
    BOOST_CHECK(parents.size() == 2 && parents[0] == v[0] && parents[1] == v[1]);
    // the CPT keeps the FILE's parent order (cpt.assign(conditional_vertexes, target), bif.hpp:104-113)
    auto const listed = v[2]->cpt.condition_node();
    BOOST_CHECK(listed.size() == 2 && listed[0] == v[1] && listed[1] == v[0]);
    bn::condition_t cond = {{v[0], 1}, {v[1], 2}};          // Cloudy = yes, Rain = heavy
    auto const row = v[2]->cpt[cond];
    BOOST_CHECK(row.first);
    BOOST_CHECK_CLOSE(row.second[0], 0.05, 1e-12);
    BOOST_CHECK_CLOSE(row.second[1], 0.95, 1e-12);
    BOOST_CHECK_CLOSE(v[0]->cpt[bn::condition_t()].second[1], 0.5, 1e-12);
}

BOOST_AUTO_TEST_CASE( bif_write_round_trip )
{
    bn::serializer::bif io;
    auto first = io.from_data(kBif);
    std::string const text = bn::serializer::bif::write(std::get<0>(first), std::get<1>(first));
    auto second = io.from_data(text);
    bn::flat_network const a = bn::flatten(std::get<0>(first)), b = bn::flatten(std::get<0>(second));
    BOOST_CHECK(a.same_topology(b));
    BOOST_CHECK(a.cpt == b.cpt);
    BOOST_CHECK(std::get<1>(second).node_name[2] == "Wet");
}

BOOST_AUTO_TEST_CASE( bif_errors_throw )
{
    bn::serializer::bif io;
    BOOST_CHECK_THROW(io.from_data("variable a { type discrete [ 2 ] { x, y }; } probability ( b ) { table 1, 0; }"), std::runtime_error);
    BOOST_CHECK_THROW(io.from_data("network a { } network b { }"), std::runtime_error);
    BOOST_CHECK_THROW(io.from_data("variable a { type discrete [ 2 ] { x }; }"), std::runtime_error);
    BOOST_CHECK_THROW(io.from_file("/nonexistent/file.bif"), std::runtime_error);
}

BOOST_AUTO_TEST_CASE( dsc_parse_like_the_reference )
{
    bn::serializer::dsc reader;
    bn::graph_t const graph = reader.from_data(kDsc);
    auto const& v = graph.vertex_list();
    BOOST_CHECK(v.size() == 2);
    BOOST_CHECK(v[1]->selectable_num == 3);
    BOOST_CHECK(graph.in_vertexes(v[1]).size() == 1 && graph.in_vertexes(v[1])[0] == v[0]);
    bn::condition_t cond = {{v[0], 1}};
    BOOST_CHECK_CLOSE(v[1]->cpt[cond].second[2], 0.3, 1e-12);
    BOOST_CHECK(reader.database().graph_name == "sprinkler");
    BOOST_CHECK(reader.database().node_name.at(0) == "Cloudy");
    std::istringstream iss(kDsc);
    BOOST_CHECK(bn::serializer::stream_to_lines(iss).size() == 19);
}
