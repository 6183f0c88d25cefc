// fastfmt.h -- the few printf conversions of the result rows without stdio: decimal integers and "%.3E".
//
// The alignment DB is text (Matcher::resultToBuffer, M/alignment/Matcher.cpp:323-370): a 1 M-spacer search prints 14 M rows,
// and glibc's "%.3E" alone costs more host time than the whole GPU pass.  put_e3 prints the same bytes as printf("%.3E"):
// the value is scaled to four significant digits in 80-bit arithmetic, and whenever the scaled value lies so close to a
// rounding boundary that the 80-bit product could be on the wrong side, it defers to snprintf (exact by construction).
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace spbhost {

static inline char *put_uint(char *p, unsigned long long v) {
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}

static inline char *put_int(char *p, long long v) {
    if (v < 0) { *p++ = '-'; return put_uint(p, (unsigned long long)(-(v + 1)) + 1ull); }
    return put_uint(p, (unsigned long long)v);
}

// 10^n as long double for n in [-350, 350]; exact for 0 <= n <= 27, otherwise within an ulp (2^-64 relative)
struct Pow10Table {
    long double v[701];
    Pow10Table() {
        v[350] = 1.0L;
        for (int i = 1; i <= 350; i++) v[350 + i] = v[350 + i - 1] * 10.0L;
        for (int i = 1; i <= 350; i++) v[350 - i] = 1.0L / v[350 + i];
    }
};
static inline long double pow10_ld(int n) {
    static const Pow10Table tab;      // (initialisation of a function-local static is thread-safe)
    return tab.v[350 + n];
}

// printf("%.3E", x) for every double; returns the end of the output
static inline char *put_e3(char *p, double x) {
    if (!(x > 0.0) || !isfinite(x) || x < 1e-300 || x > 1e300) {      // zero, negative, NaN / Inf, the far ends: the library
        return p + sprintf(p, "%.3E", x);
    }
    int e2 = 0;
    frexp(x, &e2);
    int e10 = (int)floor((e2 - 1) * 0.30102999566398120);          // floor(log10(x)) or one less
    long double m = (long double)x * pow10_ld(3 - e10);             // in [1000, 100000)
    if (m >= 10000.0L) { e10++; m = (long double)x * pow10_ld(3 - e10); }
    if (m < 1000.0L) { e10--; m = (long double)x * pow10_ld(3 - e10); }
    const long double fl = floorl(m), frac = m - fl;
    // a tie or anything within the error of the 80-bit product (a few 2^-64 relative, i.e. < 1e-14 absolute here): the library decides
    if (fabsl(frac - 0.5L) < 1e-9L || m < 1000.0L || m >= 10000.0L + 1.0L) return p + sprintf(p, "%.3E", x);
    unsigned d = (unsigned)fl + (frac > 0.5L ? 1u : 0u);
    if (d >= 10000u) { d = 1000u; e10++; }
    p[0] = (char)('0' + d / 1000u);
    p[1] = '.';
    p[2] = (char)('0' + d / 100u % 10u);
    p[3] = (char)('0' + d / 10u % 10u);
    p[4] = (char)('0' + d % 10u);
    p[5] = 'E';
    p[6] = e10 < 0 ? '-' : '+';
    const unsigned ae = (unsigned)(e10 < 0 ? -e10 : e10);
    p += 7;
    if (ae >= 100u) { *p++ = (char)('0' + ae / 100u); }
    *p++ = (char)('0' + ae / 10u % 10u);
    *p++ = (char)('0' + ae % 10u);
    return p;
}

}  // namespace spbhost
