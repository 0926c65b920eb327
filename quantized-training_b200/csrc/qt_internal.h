// qt_internal.h -- shared by the translation units of libqt_b200.so (not installed).
#pragma once
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/qt_b200.h"
#include "qt_round.h"

void qt_set_error(const char *fmt, ...);
// qt_format_t -> kernel parameters; QT_ERR_INVALID_ARGUMENT if the struct is inconsistent
int qt_make_round(const qt_format_t *fmt, QtRound *P);

struct QtLutCfg;
// which switches the binade-constant path needs for this format; QT_NO_LUT if it has no such path
int qt_lut_config(const QtRound &P, QtLutCfg *cfg);
// 1 when round_fmt(u) of a bf16 u with 0 < |u| < 2^-120 depends on the sign of u only (true for every format whose
// smallest magnitudes / flush thresholds lie above 2^-120): a quotient in that range may then be computed inexactly
// as long as its sign and zero-ness are right.
int qt_tiny_safe(const QtRound &P);
