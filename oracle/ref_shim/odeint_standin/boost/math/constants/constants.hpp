// Stand-in for <boost/math/constants/constants.hpp> (Boost is not installed): the one constant the reference's
// util/angles.h uses.  boost::math::double_constants::pi is the double nearest to pi, as here.
#pragma once
namespace boost { namespace math { namespace double_constants {
constexpr double pi = 3.141592653589793238462643383279502884;
}}}  // namespace boost::math::double_constants
