// supersonic/expression/core/comparison_bound_expressions.h -- the reference's include path; the declarations live in supersonic/expression.h.
#ifndef SUPERSONIC_B200_HOST_EXPRESSION_CORE_COMPARISON_BOUND_EXPRESSIONS_H_
#define SUPERSONIC_B200_HOST_EXPRESSION_CORE_COMPARISON_BOUND_EXPRESSIONS_H_
#include "supersonic/expression.h"
#endif
