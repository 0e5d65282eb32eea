// bayesian/graph.hpp — bn::graph_t / bn::vertex_t / bn::cpt_t and the topology flattening pass.
//
// Drop-in for the reference's bayesian/graph.hpp (godai0519/BayesianNetwork, graph.hpp:12-529) on
// the belief-propagation path: same types, same member functions, same observable results
// (neighbour lists in ascending vertex-list order, add_edge refusing cycles, cpt lookups
// independent of the order a condition map was filled in).  Written from scratch around the data
// the GPU path needs instead of around a dense N x N shared_ptr matrix:
//
//   reference (graph.hpp)                          here
//   ---------------------------------------------  -------------------------------------------
//   adjacency  N x N shared_ptr<edge_t>   :483-485  per-vertex sorted neighbour lists, O(deg) queries
//   source/target: O(N^2) scan per edge   :416-481  hash map edge -> (from, to)
//   is_able_trace: DFS without visited set :437-449  DFS with a visited set (same answers)
//   cpt_t: unordered_map<condition_t, row> :57-154   dense row-major table, mixed-radix index
//                                                    (first parent slowest: the enumeration order
//                                                    of all_combination_pattern,
//                                                    belief_propagation.hpp:269-295)
//   cpt[cond] on a miss: dangling reference :117-124  first == false and an empty row (defined)
//
// New on top of the reference's surface (used by inference/belief_propagation.hpp):
//   bn::flat_network / bn::flatten(graph)  — CSR topology + CPT arena in the layout of
//   bnbp_flat_network (include/bnbp.h), i.e. what the reference re-derives from hash maps on every
//   sweep (belief_propagation.hpp:78-101 -> graph.hpp:362-481, :117-147), computed once.
#ifndef BNB200_BAYESIAN_GRAPH_HPP
#define BNB200_BAYESIAN_GRAPH_HPP

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace bn {

struct vertex_t;
struct edge_t;
typedef std::shared_ptr<vertex_t> vertex_type;
typedef std::shared_ptr<edge_t> edge_type;
typedef std::unordered_map<vertex_type, int> condition_t;   // vertex -> selected state

} // namespace bn

namespace std {

// condition_t is used as a hash-map key (cpt_t::table_type).  The value must not depend on the
// iteration order of the map, so the per-entry hashes are combined with a commutative sum.
template <> struct hash<bn::condition_t> {
    std::size_t operator()(bn::condition_t const& cond) const noexcept
    {
        std::size_t total = 0x9e3779b97f4a7c15ull ^ cond.size();
        for (auto const& entry : cond) {
            std::size_t h = std::hash<bn::vertex_type>()(entry.first);
            h ^= static_cast<std::size_t>(entry.second) + 0x9e3779b9u + (h << 6) + (h >> 2);
            total += h * 0xff51afd7ed558ccdull;
        }
        return total;
    }
};

} // namespace std

namespace bn {

// Conditional probability table of one vertex: one row of `selectable_num` probabilities per
// configuration of the parent vertices.
class cpt_t {
public:
    typedef std::unordered_map<condition_t, std::vector<double>> table_type;

    cpt_t() = default;
    explicit cpt_t(std::vector<vertex_type> const& parent_nodes, vertex_type const& target_node)
    {
        assign(parent_nodes, target_node);
    }

    // (Re)shape the table for these parents; every row is zero-filled (graph.hpp:490-525).
    inline void assign(std::vector<vertex_type> const& parent_nodes, vertex_type const& target_node);

    // Rows whose configuration agrees with every entry of `cond` (an empty cond selects all rows).
    table_type filter(condition_t const& cond) const
    {
        table_type out;
        for (std::size_t q = 0; q < rows_.size(); ++q) {
            condition_t const key = configuration(q);
            bool keep = true;
            for (auto const& want : cond) {
                auto const it = key.find(want.first);
                if (it == key.end() || it->second != want.second) { keep = false; break; }
            }
            if (keep) out.emplace(key, rows_[q]);
        }
        return out;
    }

    std::vector<vertex_type> condition_node() const { return parents_; }

    // every configuration operator[] accepts
    std::vector<condition_t> pattern() const
    {
        std::vector<condition_t> out;
        out.reserve(rows_.size());
        for (std::size_t q = 0; q < rows_.size(); ++q) out.push_back(configuration(q));
        return out;
    }

    // Row of the configuration that matches `cond` exactly.  first == false on a miss; second is
    // then an empty scratch row (the reference leaves it dangling, graph.hpp:120-124).
    std::pair<bool, std::vector<double>&> operator[](condition_t const& cond)
    {
        std::size_t q = 0;
        if (!locate(cond, q)) {
            miss_.clear();
            return std::pair<bool, std::vector<double>&>(false, miss_);
        }
        return std::pair<bool, std::vector<double>&>(true, rows_[q]);
    }
    std::pair<bool, std::vector<double> const&> operator[](condition_t const& cond) const
    {
        std::size_t q = 0;
        if (!locate(cond, q)) return std::pair<bool, std::vector<double> const&>(false, empty_row());
        return std::pair<bool, std::vector<double> const&>(true, rows_[q]);
    }

    // ---- dense access (new; used by bn::flatten) ------------------------------------------------
    std::size_t row_count() const { return rows_.size(); }
    std::vector<double> const& row(std::size_t const q) const { return rows_[q]; }
    std::vector<std::size_t> const& radix() const { return radix_; }

    // configuration index of `cond` in mixed radix over condition_node(), first parent slowest
    bool locate(condition_t const& cond, std::size_t& q) const
    {
        if (rows_.empty() || cond.size() != parents_.size()) return false;
        std::size_t index = 0;
        for (std::size_t j = 0; j < parents_.size(); ++j) {
            auto const it = cond.find(parents_[j]);
            if (it == cond.end() || it->second < 0 || static_cast<std::size_t>(it->second) >= radix_[j]) return false;
            index = index * radix_[j] + static_cast<std::size_t>(it->second);
        }
        q = index;
        return true;
    }

private:
    condition_t configuration(std::size_t q) const
    {
        condition_t key;
        for (std::size_t j = parents_.size(); j-- > 0;) {
            key[parents_[j]] = static_cast<int>(q % radix_[j]);
            q /= radix_[j];
        }
        return key;
    }
    static std::vector<double> const& empty_row()
    {
        static std::vector<double> const none;
        return none;
    }

    std::vector<vertex_type> parents_;
    std::vector<std::size_t> radix_;              // selectable_num of each parent at assign() time
    std::vector<std::vector<double>> rows_;       // [configuration][state of the target]
    std::vector<double> miss_;
};

struct vertex_t {
    int id = 0;
    std::size_t selectable_num = 0;   // number of states
    cpt_t cpt;
};

struct edge_t {};   // identity only

struct database_t {
    std::string graph_name;
    std::unordered_map<std::size_t, std::string> node_name;
    std::unordered_map<std::size_t, std::vector<std::string>> options_name;
};

inline void cpt_t::assign(std::vector<vertex_type> const& parent_nodes, vertex_type const& target_node)
{
    std::vector<std::size_t> radix;
    std::size_t configurations = 1;
    for (vertex_type const& p : parent_nodes) {
        radix.push_back(p->selectable_num);
        configurations *= p->selectable_num;
    }
    std::vector<std::vector<double>> rows(configurations, std::vector<double>(target_node->selectable_num, 0.0));
    parents_ = parent_nodes;
    radix_.swap(radix);
    rows_.swap(rows);
}

// Directed acyclic graph over shared vertex / edge objects.
class graph_t {
public:
    graph_t() = default;
    virtual ~graph_t() = default;
    graph_t(graph_t const&) = default;            // shallow like the reference: vertices are shared
    graph_t(graph_t&& other) noexcept { swap(other); }
    graph_t& operator=(graph_t const& rhs)
    {
        graph_t(rhs).swap(*this);
        return *this;
    }
    graph_t& operator=(graph_t&& rhs) noexcept
    {
        swap(rhs);
        return *this;
    }

    void swap(graph_t& other) noexcept
    {
        vertex_list_.swap(other.vertex_list_);
        edge_list_.swap(other.edge_list_);
        position_.swap(other.position_);
        ends_.swap(other.ends_);
        out_.swap(other.out_);
        in_.swap(other.in_);
    }
    friend inline void swap(graph_t& lhs, graph_t& rhs) noexcept { lhs.swap(rhs); }

    std::vector<vertex_type> const& vertex_list() const { return vertex_list_; }
    std::vector<edge_type> const& edge_list() const { return edge_list_; }

    // Deep copy: new vertex objects (id, selectable_num, CPT) and new edges.  Unlike the reference
    // (graph.hpp:224-247, whose cloned CPTs stay keyed by the ORIGINAL vertices) the cloned CPTs
    // are re-keyed to the cloned parents, so a clone can be used for inference.
    graph_t clone() const
    {
        graph_t copy;
        for (std::size_t i = 0; i < vertex_list_.size(); ++i) {
            vertex_type v = copy.add_vertex();
            v->id = vertex_list_[i]->id;
            v->selectable_num = vertex_list_[i]->selectable_num;
        }
        for (edge_type const& e : edge_list_) {
            auto const where = ends_.find(e.get());
            if (where != ends_.end())
                copy.add_edge(copy.vertex_list_[where->second.first], copy.vertex_list_[where->second.second]);
        }
        for (std::size_t i = 0; i < vertex_list_.size(); ++i) {
            cpt_t const& src = vertex_list_[i]->cpt;
            std::vector<vertex_type> parents;
            bool inside = true;
            for (vertex_type const& p : src.condition_node()) {
                auto const at = position_.find(p.get());
                if (at == position_.end()) { inside = false; break; }
                parents.push_back(copy.vertex_list_[at->second]);
            }
            if (!inside || src.row_count() == 0) continue;
            cpt_t& dst = copy.vertex_list_[i]->cpt;
            dst.assign(parents, copy.vertex_list_[i]);
            if (dst.row_count() != src.row_count()) continue;
            std::vector<condition_t> const keys = dst.pattern();
            for (condition_t const& key : keys) {
                std::size_t q = 0;
                if (dst.locate(key, q)) dst[key].second = src.row(q);
            }
        }
        return copy;
    }

    vertex_type add_vertex()
    {
        vertex_type v = std::make_shared<vertex_t>();
        position_[v.get()] = vertex_list_.size();
        vertex_list_.push_back(v);
        out_.emplace_back();
        in_.emplace_back();
        return v;
    }

    // nullptr when the edge would close a directed cycle, when an end is not a vertex of this
    // graph, or when the edge exists already (graph.hpp:268-291)
    edge_type add_edge(vertex_type const& from, vertex_type const& to)
    {
        std::size_t a = 0, b = 0;
        if (!find_index(from, a) || !find_index(to, b)) return nullptr;
        if (reaches(b, a)) return nullptr;
        if (linked(a, b)) return nullptr;
        edge_type e = std::make_shared<edge_t>();
        edge_list_.push_back(e);
        ends_[e.get()] = std::make_pair(a, b);
        insert_sorted(out_[a], link{b, e});
        insert_sorted(in_[b], link{a, e});
        return e;
    }

    // Removes the vertex AND its incident edges (the reference leaves them behind in edge_list(),
    // graph.hpp:295-304, where source()/target() can no longer resolve them).
    bool erase_vertex(vertex_type const& v)
    {
        std::size_t gone = 0;
        if (!find_index(v, gone)) return false;
        std::vector<edge_type> incident;
        for (link const& l : out_[gone]) incident.push_back(l.edge);
        for (link const& l : in_[gone]) incident.push_back(l.edge);
        for (edge_type const& e : incident) erase_edge(e);
        vertex_list_.erase(vertex_list_.begin() + static_cast<std::ptrdiff_t>(gone));
        out_.erase(out_.begin() + static_cast<std::ptrdiff_t>(gone));
        in_.erase(in_.begin() + static_cast<std::ptrdiff_t>(gone));
        position_.erase(v.get());
        for (auto& p : position_)
            if (p.second > gone) --p.second;
        for (auto& e : ends_) {
            if (e.second.first > gone) --e.second.first;
            if (e.second.second > gone) --e.second.second;
        }
        for (auto* side : {&out_, &in_})
            for (auto& list : *side)
                for (link& l : list)
                    if (l.other > gone) --l.other;
        return true;
    }

    bool erase_edge(edge_type const& e)
    {
        auto const where = e ? ends_.find(e.get()) : ends_.end();
        if (where == ends_.end()) return false;
        std::size_t const a = where->second.first, b = where->second.second;
        remove_link(out_[a], b);
        remove_link(in_[b], a);
        ends_.erase(where);
        edge_list_.erase(std::remove(edge_list_.begin(), edge_list_.end(), e), edge_list_.end());
        return true;
    }

    bool erase_all_vertex()
    {
        vertex_list_.clear();
        edge_list_.clear();
        position_.clear();
        ends_.clear();
        out_.clear();
        in_.clear();
        return true;
    }

    bool erase_all_edge()
    {
        edge_list_.clear();
        ends_.clear();
        for (auto& l : out_) l.clear();
        for (auto& l : in_) l.clear();
        return true;
    }

    // Reverse an edge; when that would close a cycle the original edge is put back (as a new edge
    // object, like the reference) and nullptr is returned (graph.hpp:338-358).
    edge_type change_edge_direction(edge_type const& e)
    {
        vertex_type const from = source(e), to = target(e);
        if (!erase_edge(e)) return nullptr;
        if (edge_type const reversed = add_edge(to, from)) return reversed;
        add_edge(from, to);
        return nullptr;
    }

    std::vector<edge_type> out_edges(vertex_type const& from) const { return edges_of(out_, from); }
    std::vector<edge_type> in_edges(vertex_type const& to) const { return edges_of(in_, to); }
    // children / parents in ascending vertex-list index (graph.hpp:362-413): the order every
    // message loop of the BP path relies on
    std::vector<vertex_type> out_vertexes(vertex_type const& from) const { return vertexes_of(out_, from); }
    std::vector<vertex_type> in_vertexes(vertex_type const& to) const { return vertexes_of(in_, to); }

    vertex_type source(edge_type const& e) const
    {
        auto const where = e ? ends_.find(e.get()) : ends_.end();
        return where == ends_.end() ? nullptr : vertex_list_[where->second.first];
    }
    vertex_type target(edge_type const& e) const
    {
        auto const where = e ? ends_.find(e.get()) : ends_.end();
        return where == ends_.end() ? nullptr : vertex_list_[where->second.second];
    }

    // Is `to` reachable from `from` along directed edges (true for from == to, graph.hpp:437-449)?
    bool is_able_trace(vertex_type const& from, vertex_type const& to) const
    {
        if (from == to) return true;
        std::size_t a = 0, b = 0;
        if (!find_index(from, a) || !find_index(to, b)) return false;
        return reaches(a, b);
    }

    // ---- index access (new; used by bn::flatten) ---------------------------------------------------
    bool find_index(vertex_type const& v, std::size_t& index) const
    {
        auto const at = v ? position_.find(v.get()) : position_.end();
        if (at == position_.end()) return false;
        index = at->second;
        return true;
    }
    std::vector<std::size_t> parent_indices(std::size_t const i) const
    {
        std::vector<std::size_t> out;
        for (link const& l : in_[i]) out.push_back(l.other);
        return out;
    }

private:
    struct link {
        std::size_t other;   // index of the vertex at the far end
        edge_type edge;
    };
    typedef std::vector<std::vector<link>> side_type;

    static void insert_sorted(std::vector<link>& list, link const& l)
    {
        auto const pos = std::lower_bound(list.begin(), list.end(), l.other,
                                          [](link const& x, std::size_t const key) { return x.other < key; });
        list.insert(pos, l);
    }
    static void remove_link(std::vector<link>& list, std::size_t const other)
    {
        list.erase(std::remove_if(list.begin(), list.end(), [other](link const& x) { return x.other == other; }),
                   list.end());
    }
    bool linked(std::size_t const a, std::size_t const b) const
    {
        for (link const& l : out_[a])
            if (l.other == b) return true;
        return false;
    }
    bool reaches(std::size_t const from, std::size_t const to) const
    {
        if (from == to) return true;
        std::vector<char> seen(vertex_list_.size(), 0);
        std::vector<std::size_t> stack(1, from);
        seen[from] = 1;
        while (!stack.empty()) {
            std::size_t const u = stack.back();
            stack.pop_back();
            for (link const& l : out_[u]) {
                if (l.other == to) return true;
                if (!seen[l.other]) { seen[l.other] = 1; stack.push_back(l.other); }
            }
        }
        return false;
    }
    std::vector<edge_type> edges_of(side_type const& side, vertex_type const& v) const
    {
        std::vector<edge_type> out;
        std::size_t i = 0;
        if (!find_index(v, i)) return out;
        for (link const& l : side[i]) out.push_back(l.edge);
        return out;
    }
    std::vector<vertex_type> vertexes_of(side_type const& side, vertex_type const& v) const
    {
        std::vector<vertex_type> out;
        std::size_t i = 0;
        if (!find_index(v, i)) return out;
        for (link const& l : side[i]) out.push_back(vertex_list_[l.other]);
        return out;
    }

    std::vector<vertex_type> vertex_list_;
    std::vector<edge_type> edge_list_;
    std::unordered_map<vertex_t const*, std::size_t> position_;                     // vertex -> index
    std::unordered_map<edge_t const*, std::pair<std::size_t, std::size_t>> ends_;   // edge -> (from, to)
    side_type out_, in_;                                                             // sorted by `other`
};

// ------------------------------------------------------------------------------------------------
// Topology flattening: the arrays of bnbp_flat_network (include/bnbp.h).
//   node i          = vertex_list()[i]
//   parents of i    = in_vertexes(vertex i), ascending index          (graph.hpp:389-413)
//   cpt of i        = row-major [configuration of in_vertexes][state], first parent slowest
//                     (all_combination_pattern, belief_propagation.hpp:269-295); every row is read
//                     through vertex->cpt[condition] exactly as calculate_pi / calculate_lambda_k
//                     do (:182-196, :245-262), so the CPT may list its parents in any order.
// The reference dereferences a missing row (UB, graph.hpp:120-124); flatten() throws instead.
struct flat_network {
    std::vector<std::int32_t> card;
    std::vector<std::int32_t> parent_off;
    std::vector<std::int32_t> parents;
    std::vector<std::int64_t> cpt_off;
    std::vector<double> cpt;

    bool same_topology(flat_network const& o) const
    {
        return card == o.card && parent_off == o.parent_off && parents == o.parents && cpt_off == o.cpt_off;
    }
};

inline flat_network flatten(graph_t const& graph)
{
    std::vector<vertex_type> const& nodes = graph.vertex_list();
    flat_network f;
    f.parent_off.push_back(0);
    f.cpt_off.push_back(0);
    for (std::size_t i = 0; i < nodes.size(); ++i) {
        vertex_type const& x = nodes[i];
        if (x->selectable_num == 0)
            throw std::invalid_argument("bn::flatten: vertex " + std::to_string(i) + " has selectable_num == 0");
        f.card.push_back(static_cast<std::int32_t>(x->selectable_num));
        std::vector<std::size_t> const pidx = graph.parent_indices(i);
        std::size_t configurations = 1;
        for (std::size_t const p : pidx) {
            f.parents.push_back(static_cast<std::int32_t>(p));
            configurations *= nodes[p]->selectable_num;
        }
        f.parent_off.push_back(static_cast<std::int32_t>(f.parents.size()));
        cpt_t const& table = x->cpt;
        condition_t cond;
        std::vector<std::size_t> state(pidx.size(), 0);
        for (std::size_t q = 0; q < configurations; ++q) {
            for (std::size_t j = 0; j < pidx.size(); ++j) cond[nodes[pidx[j]]] = static_cast<int>(state[j]);
            auto const found = table[cond];
            if (!found.first)
                throw std::invalid_argument("bn::flatten: vertex " + std::to_string(i) +
                                            " has no CPT row for a configuration of its parents");
            if (found.second.size() != x->selectable_num)
                throw std::invalid_argument("bn::flatten: vertex " + std::to_string(i) +
                                            " has a CPT row whose length differs from selectable_num");
            f.cpt.insert(f.cpt.end(), found.second.begin(), found.second.end());
            for (std::size_t j = pidx.size(); j-- > 0;) {           // last parent fastest
                if (++state[j] < nodes[pidx[j]]->selectable_num) break;
                state[j] = 0;
            }
        }
        f.cpt_off.push_back(static_cast<std::int64_t>(f.cpt.size()));
    }
    return f;
}

} // namespace bn

#endif // BNB200_BAYESIAN_GRAPH_HPP
