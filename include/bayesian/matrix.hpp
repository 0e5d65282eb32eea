// bayesian/matrix.hpp — bn::matrix_type, the host value type of the BP path.
//
// Drop-in for the reference's bayesian/matrix.hpp (godai0519/BayesianNetwork, matrix.hpp:10-161):
// same members, same semantics, written from scratch.  On the belief-propagation path a
// matrix_type is always a 1 x r row (evidence rows going in, belief rows coming out,
// belief_propagation.hpp:31,151-158); the device-side counterpart of the reference's per-message
// matrices is the batch-minor state arena inside libbnbp, not this class.
//
// Contract kept from the reference:
//   * operator[](row) hands out the row as std::vector<double>& -- the reference tests assign
//     whole rows through it (libs/bayesian/test/belief_propagation.cpp:193-194);
//   * resize(h, w, fill) keeps existing entries and fills new ones (matrix.hpp:23-35);
//   * assign(first, last) copies h*w values row-major and reports false when the range is too
//     short (matrix.hpp:38-58);
//   * operator% / %= element-wise product, operator* / *= matrix product, scalar * on either side;
//   * `matrix_type` is also visible unqualified (matrix.hpp:134-136): the reference tests use it so.
#ifndef BNB200_BAYESIAN_MATRIX_HPP
#define BNB200_BAYESIAN_MATRIX_HPP

#include <cassert>
#include <cstddef>
#include <initializer_list>
#include <iterator>
#include <utility>
#include <vector>

namespace bn {

class matrix_type {
public:
    typedef std::vector<double> row_type;

    matrix_type() = default;
    matrix_type(std::size_t const height, std::size_t const width, double const default_value = 0.0)
        : rows_(height, row_type(width, default_value)), width_(width)
    {
    }
    virtual ~matrix_type() = default;

    std::size_t height() const { return rows_.size(); }
    std::size_t width() const { return width_; }

    void resize(std::size_t const height, std::size_t const width, double const default_value = 0.0)
    {
        rows_.resize(height);
        for (row_type& r : rows_) r.resize(width, default_value);
        width_ = width;
    }

    template <class InputIterator>
    bool assign(InputIterator first, InputIterator const& last)
    {
        auto const available = std::distance(first, last);
        if (available < 0 || static_cast<std::size_t>(available) < width_ * rows_.size()) return false;
        for (row_type& r : rows_)
            for (double& cell : r) cell = *first++;
        return true;
    }

    row_type& operator[](std::size_t const row) { return rows_[row]; }
    row_type const& operator[](std::size_t const row) const { return rows_[row]; }

    // element-wise (Hadamard) product
    matrix_type& operator%=(matrix_type const& rhs)
    {
        assert(width() == rhs.width() && height() == rhs.height());
        for (std::size_t y = 0; y < rows_.size(); ++y)
            for (std::size_t x = 0; x < width_; ++x) rows_[y][x] *= rhs.rows_[y][x];
        return *this;
    }
    matrix_type operator%(matrix_type const& rhs) const
    {
        matrix_type out(*this);
        out %= rhs;
        return out;
    }

    // matrix product (the reference returns *= by value, matrix.hpp:96; kept)
    matrix_type operator*=(matrix_type const& rhs)
    {
        assert(width() == rhs.height());
        std::vector<row_type> out(height(), row_type(rhs.width(), 0.0));
        for (std::size_t y = 0; y < height(); ++y)
            for (std::size_t k = 0; k < rhs.height(); ++k) {
                double const a = rows_[y][k];
                for (std::size_t x = 0; x < rhs.width(); ++x) out[y][x] += a * rhs.rows_[k][x];
            }
        rows_.swap(out);
        width_ = rhs.width();
        return *this;
    }
    matrix_type operator*(matrix_type const& rhs) const
    {
        matrix_type out(*this);
        out *= rhs;
        return out;
    }

    // ---- additions used by the batched front ends (not in the reference) ------------------------------
    // a 1 x n row from a list of values: the shape of every evidence / belief matrix on the BP path
    static matrix_type row(std::initializer_list<double> values)
    {
        matrix_type m(1, values.size());
        std::size_t i = 0;
        for (double const v : values) m.rows_[0][i++] = v;
        return m;
    }

    // row-major copy of all entries: what the flat C ABI (bnbp_evidence::ev_values) takes
    std::vector<double> flat() const
    {
        std::vector<double> out;
        out.reserve(rows_.size() * width_);
        for (row_type const& r : rows_) out.insert(out.end(), r.begin(), r.end());
        return out;
    }

    bool same_shape(matrix_type const& other) const { return height() == other.height() && width() == other.width(); }

    double sum() const
    {
        double total = 0.0;
        for (row_type const& r : rows_)
            for (double const cell : r) total += cell;
        return total;
    }

    // every entry divided by the plain sum, no zero guard: the normalisation of the BP path
    // (belief_propagation.hpp:298-311; 0/0 stays NaN exactly as there)
    matrix_type normalized() const
    {
        matrix_type out(*this);
        double const total = sum();
        for (row_type& r : out.rows_)
            for (double& cell : r) cell /= total;
        return out;
    }

private:
    std::vector<row_type> rows_;
    std::size_t width_ = 0;
};

} // namespace bn

namespace {
using bn::matrix_type;
}

template <class Scalar>
matrix_type operator*(matrix_type const& m, Scalar const& s)
{
    matrix_type out(m);
    for (std::size_t y = 0; y < out.height(); ++y)
        for (double& cell : out[y]) cell *= s;
    return out;
}

template <class Scalar>
matrix_type operator*(Scalar const& s, matrix_type const& m)
{
    return m * s;
}

#endif // BNB200_BAYESIAN_MATRIX_HPP
