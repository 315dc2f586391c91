// supersonic/expression/core/arithmetic_bound_expressions.h -- the reference's include path; the declarations live in supersonic/expression.h.
#ifndef SUPERSONIC_B200_HOST_EXPRESSION_CORE_ARITHMETIC_BOUND_EXPRESSIONS_H_
#define SUPERSONIC_B200_HOST_EXPRESSION_CORE_ARITHMETIC_BOUND_EXPRESSIONS_H_
#include "supersonic/expression.h"
#endif
