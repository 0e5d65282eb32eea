// Stand-in for <boost/optional.hpp> (Boost is not in this image) — TEST INFRASTRUCTURE.
// bayesian/sampler.hpp of the reference includes the header but never names boost::optional; this minimal
// value-or-nothing holder only lets the reference file compile where it lies (oracle/Makefile `ref`).
#ifndef BNB200_TESTS_BOOST_OPTIONAL_HPP
#define BNB200_TESTS_BOOST_OPTIONAL_HPP
namespace boost {
struct none_t {};
static const none_t none = none_t();
template <class T> class optional {
public:
    optional() : has_(false), value_() {}
    optional(none_t) : has_(false), value_() {}
    optional(T const& v) : has_(true), value_(v) {}
    explicit operator bool() const { return has_; }
    T& operator*() { return value_; }
    T const& operator*() const { return value_; }
    T* operator->() { return &value_; }
    T const* operator->() const { return &value_; }
    T const& get() const { return value_; }
private:
    bool has_;
    T value_;
};
} // namespace boost
#endif
