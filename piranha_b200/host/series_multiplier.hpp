// The plugin point of the multiplier path, as in the reference (include/piranha/series_multiplier.hpp:37-56): an empty
// primary template; a specialisation is constructed from two series (whose symbol sets have already been merged) and
// its operator()() returns the product.  The specialisation for polynomial<integer, kronecker_monomial<int64_t>> that
// runs on the GPU lives in polynomial.hpp, next to the series type, like the reference's (polynomial.hpp:970-992).
#ifndef PB200_SERIES_MULTIPLIER_HPP
#define PB200_SERIES_MULTIPLIER_HPP

namespace pb200
{

template <typename Series, typename Enable = void>
class series_multiplier
{
};

} // namespace pb200

#endif
