// format_dispatch.cuh — maps a runtime (exp, man, rounding) to a compile-time
// specialised format policy for the common formats and to the runtime policy
// otherwise.  Usage: CPD_DISPATCH_FORMAT(e, m, rounding, FMT, { kernel<decltype(FMT)>(..., FMT); })
#pragma once
#include "../common/fpformat.h"

#define CPD_DISPATCH_FORMAT(E, M, ROUNDING, NAME, ...)                                   \
  do {                                                                                   \
    const int _e = (E), _m = (M), _r = (ROUNDING);                                       \
    if (_r == ::cpd::ROUND_REF && _e == 4 && _m == 3) { ::cpd::FmtT<4, 3> NAME; __VA_ARGS__ } \
    else if (_r == ::cpd::ROUND_REF && _e == 5 && _m == 2) { ::cpd::FmtT<5, 2> NAME; __VA_ARGS__ } \
    else if (_r == ::cpd::ROUND_REF && _e == 5 && _m == 10) { ::cpd::FmtT<5, 10> NAME; __VA_ARGS__ } \
    else if (_r == ::cpd::ROUND_REF && _e == 8 && _m == 7) { ::cpd::FmtT<8, 7> NAME; __VA_ARGS__ } \
    else if (_r == ::cpd::ROUND_REF && _e == 8 && _m == 23) { ::cpd::FmtT<8, 23> NAME; __VA_ARGS__ } \
    else { ::cpd::FmtR NAME{::cpd::make_format(_e, _m, _r)}; __VA_ARGS__ }              \
  } while (0)
