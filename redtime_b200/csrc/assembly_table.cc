// Linear assembly of the mode-coupling outputs from the bilinear quadratures.
//
//   A_{acd,bef}(k) (14 unique), R^ell_{abc}(k) (24), P_{T,jm}(k) (9), P_{MR,n}(k) (8)
//   = sum_terms coef * k^kpow * {J | PZ | Jn0}[9 n + 3 ab + cd](k)
//
// This is the physics of redTime.cc:813-1279 (1-loop vertex algebra of Upadhye 2019 /
// TNS 2010 / McDonald-Roy 2009) re-expressed as a sparse coefficient table that the
// device assembly kernel walks.  Recurring combinations are factored into "forms"; the
// table is verified term-for-term against the reference in tests/test_tables.py (the
// reference's own J/PZ arrays pushed through this table must reproduce its A/R/PT/PMR).
#include "fastpt_tables.h"

#include <cmath>

namespace rtrg {

namespace {

struct Builder {
  std::vector<AsmTerm> t;
  int row = 0, kpow = 0;
  double scale = 1.0;
  void add(int src, int n, int pair, double c, int extra_kpow = 0) {
    AsmTerm a;
    a.row = (short)row;
    a.src = (short)src;
    a.index = (short)(9 * n + pair);
    a.kpow = (short)(kpow + extra_kpow);
    a.coef = c * scale;
    t.push_back(a);
  }
  void J(int n, int pair, double c) { add(0, n, pair, c); }
  void PZ(int n, int pair, double c) { add(1, n, pair, c); }
  void J0(int n, int pair, double c, int kp) { add(2, n, pair, c, kp); }  // Jn0 / k^(-kp)
  void Jlo(double c) { add(3, 0, 0, c); }
};

// A_{001,bef}-type J combination (redTime.cc:820-823 and siblings)
void formA1(Builder &B, int x, int y, int z, double s = 1.0) {
  B.J(4, x, s / 6), B.J(2, x, s / 2), B.J(0, x, s / 4), B.J(1, x, s / 12);
  B.J(3, y, s / 6), B.J(2, y, s / 4), B.J(2, z, s / 4), B.J(0, y, s / 3);
}
// A_{001,1ef} / A_{111,0ef}-type J combination (redTime.cc:858-861)
void formA2(Builder &B, int x, int y, double s = 1.0) {
  B.J(5, x, s / 5), B.J(3, x, s / 2), B.J(4, x, s / 6), B.J(2, x, s * 0.55);
  B.J(2, y, s / 4), B.J(0, x, s / 4), B.J(1, x, s / 12);
}
// A_{111,1ef}-type J combination (redTime.cc:930-936)
void formA3(Builder &B, int x, int y) {
  B.J(6, x, 8.0 / 35), B.J(5, x, 0.4), B.J(5, y, 0.4), B.J(3, x, 19.0 / 21);
  B.J(4, x, 1.0 / 6), B.J(4, y, 1.0 / 6), B.J(2, x, 0.6), B.J(2, y, 0.6);
  B.J(0, x, 11.0 / 30), B.J(1, x, 1.0 / 12), B.J(1, y, 1.0 / 12);
}
// P13-type pieces (redTime.cc:824-829 and :901-905)
void formPZ1(Builder &B, int p, int q) {
  B.PZ(0, p, -1.0 / 12);
  B.PZ(4, q, 1.0 / 16), B.PZ(2, q, -1.0 / 16), B.PZ(0, q, 1.0 / 16), B.PZ(1, q, 0.5 / 16);
  B.PZ(3, p, -1.0 / 16), B.PZ(1, p, 1.0 / 16), B.PZ(0, p, 3.0 / 16), B.PZ(2, p, -0.5 / 16);
}
void formPZ2(Builder &B, int p, int q, double s) {
  B.PZ(4, p, -2 * s / 16), B.PZ(2, p, 2 * s / 16), B.PZ(0, p, -2 * s / 16), B.PZ(1, p, -s / 16);
  B.PZ(6, q, 2 * s / 16), B.PZ(4, q, -4 * s / 16), B.PZ(2, q, s / 16);
}

std::vector<AsmTerm> build() {
  Builder B;
  // ---------------- A_{acd,bef}: pre_A = k/(4 pi) (redTime.cc:815) --------------------
  B.kpow = 1;
  B.scale = 1.0 / (4.0 * M_PI);
  B.row = ASM_A0 + 0, formA1(B, 1, 3, 1), formPZ1(B, 1, 3);        // A[8]
  B.row = ASM_A0 + 1, formA1(B, 2, 4, 4);                          // A[9]
  B.row = ASM_A0 + 2, formA1(B, 4, 6, 2), formPZ1(B, 4, 6);        // A[10]
  B.row = ASM_A0 + 3, formA1(B, 5, 7, 5);                          // A[11]
  B.row = ASM_A0 + 4, formA2(B, 4, 4), formPZ1(B, 2, 4);           // A[12]
  B.row = ASM_A0 + 5, formA2(B, 5, 7);                             // A[13]
  B.row = ASM_A0 + 6, formA2(B, 7, 5), formPZ1(B, 5, 7);           // A[14]
  B.row = ASM_A0 + 7, formA2(B, 8, 8);                             // A[15]
  B.row = ASM_A0 + 8, formA2(B, 1, 3, 2.0), formPZ2(B, 1, 3, 1.0); // A[56]
  B.row = ASM_A0 + 9, formA2(B, 2, 6), formA2(B, 4, 4), formPZ2(B, 4, 6, 0.5);  // A[57]
  B.row = ASM_A0 + 10, formA2(B, 5, 7, 2.0);                       // A[59]
  B.row = ASM_A0 + 11, formA3(B, 4, 4), formPZ2(B, 2, 4, 1.0);     // A[60]
  B.row = ASM_A0 + 12, formA3(B, 5, 7), formPZ2(B, 5, 7, 0.5);     // A[61]
  B.row = ASM_A0 + 13, formA3(B, 8, 8);                            // A[63]

  // ---------------- R^ell_{abc}: pre_R = 1/(2 pi k) (redTime.cc:816) ------------------
  B.kpow = -1;
  B.scale = 1.0 / (2.0 * M_PI);
  for (int a = 0; a < 2; a++)
    for (int b = 0; b < 2; b++)
      for (int c = 0; c < 2; c++) {
        const int e = 3 * b + c, t = 3 * c + b, g = 3 * c + a, h = 3 * b + a;
        // ell = 1 (redTime.cc:985-1042)
        B.row = ASM_R0 + 0 + 4 * a + 2 * b + c;
        if (a == 0) {
          B.J(5, e + 1, 0.4), B.J(2, e + 1, -1.4), B.J(2, t + 3, -1.0), B.J(0, e + 1, -2.0);
          B.J(5, t + 1, 0.4), B.J(3, e + 3, 2.0 / 3), B.J(4, t + 1, -2.0 / 3);
          B.J(2, t + 1, -2.4), B.J(0, e + 3, -5.0 / 3), B.J(1, t + 1, -1.0 / 3);
        } else {
          B.J(6, e + 4, 16.0 / 35), B.J(5, t + 4, -0.4), B.J(5, e + 4, 0.4);
          B.J(3, e + 4, -46.0 / 21), B.J(4, e + 4, -2.0 / 3), B.J(2, t + 4, -2.6);
          B.J(2, e + 4, -1.4), B.J(0, e + 4, -19.0 / 15), B.J(1, t + 4, -1.0 / 3);
        }
        if (b == 0) {
          B.PZ(0, g + 1, -13.0 / 12), B.PZ(2, g + 1, 5.0 / 16), B.PZ(1, g + 1, -7.0 / 16);
          B.PZ(4, g + 1, -0.125), B.PZ(3, g + 1, 0.375), B.PZ(0, g + 3, -0.375);
          B.PZ(2, g + 3, 7.0 / 16), B.PZ(1, g + 3, -3.0 / 16), B.PZ(4, g + 3, -0.625);
          B.PZ(6, g + 3, 0.125);
        } else {
          B.PZ(0, g + 4, -1.0 / 3);
        }
        if (c == 0) {
          B.PZ(6, h + 3, 0.125), B.PZ(4, h + 3, -0.375), B.PZ(2, h + 3, 3.0 / 16);
          B.PZ(1, h + 3, -1.0 / 16), B.PZ(0, h + 3, -0.125), B.PZ(4, h + 1, -0.125);
          B.PZ(2, h + 1, 3.0 / 16), B.PZ(1, h + 1, -3.0 / 16), B.PZ(3, h + 1, 0.125);
        } else {
          B.PZ(0, h + 4, 1.0 / 3);
        }
        // ell = 2 (redTime.cc:1044-1098)
        B.row = ASM_R0 + 8 + 4 * a + 2 * b + c;
        if (a == 0) {
          B.J(5, e + 1, 0.6), B.J(3, e + 1, 1.0), B.J(2, e + 1, -0.6), B.J(0, e + 1, -1.0);
          B.J(5, t + 1, 0.6), B.J(3, e + 3, 1.0), B.J(2, t + 1, -0.6), B.J(0, e + 3, -1.0);
        } else {
          B.J(6, e + 4, 24.0 / 35), B.J(5, t + 4, -1.0), B.J(5, e + 4, 2.2);
          B.J(3, e + 4, -2.0 / 7), B.J(2, e + 4, -0.6), B.J(2, t + 4, -0.6), B.J(0, e + 4, -0.4);
        }
        if (b == 0) {
          B.PZ(0, g + 1, -0.5), B.PZ(2, g + 1, 9.0 / 32), B.PZ(1, g + 1, -9.0 / 32);
          B.PZ(4, g + 1, -3.0 / 16), B.PZ(3, g + 1, 3.0 / 16), B.PZ(0, g + 3, -3.0 / 16);
          B.PZ(1, g + 3, -3.0 / 32), B.PZ(2, g + 3, 9.0 / 32), B.PZ(4, g + 3, -9.0 / 16);
          B.PZ(6, g + 3, 3.0 / 16);
        }
        if (c == 0) {
          B.PZ(6, h + 3, 3.0 / 16), B.PZ(4, h + 3, -9.0 / 16), B.PZ(2, h + 3, 9.0 / 32);
          B.PZ(1, h + 3, -3.0 / 32), B.PZ(0, h + 3, -3.0 / 16), B.PZ(3, h + 1, 3.0 / 16);
          B.PZ(4, h + 1, -3.0 / 16), B.PZ(1, h + 1, -9.0 / 32), B.PZ(2, h + 1, 9.0 / 32);
          B.PZ(0, h + 1, -0.5);
        }
        // ell = 3 (redTime.cc:1100-1158)
        B.row = ASM_R0 + 16 + 4 * a + 2 * b + c;
        if (a == 0) {
          B.J0(2, t + 3, 4.0 / 7, -2), B.J0(1, t + 3, -40.0 / 21, -2), B.J0(0, t + 3, 4.0 / 3, -2);
          B.J0(2, e + 3, -4.0 / 7, -2), B.J0(1, e + 3, 40.0 / 21, -2), B.J0(0, e + 3, -4.0 / 3, -2);
          B.J(5, e + 1, -1.0), B.J(2, e + 1, 1.0), B.J(3, e + 3, -5.0 / 3), B.J(0, e + 3, 5.0 / 3);
        } else {
          B.J(6, e + 4, -4.0 / 7), B.J(5, e + 4, -1.0), B.J(3, e + 4, 5.0 / 21);
          B.J(2, e + 4, 1.0), B.J(0, e + 4, 1.0 / 3);
        }
        if (b == 0) {
          B.PZ(0, g + 1, 35.0 / 32), B.PZ(5, g + 1, 5.0 / 32), B.PZ(3, g + 1, -5.0 / 8);
          B.PZ(4, g + 1, 5.0 / 32), B.PZ(2, g + 1, -5.0 / 16), B.PZ(1, g + 1, 15.0 / 32);
          B.PZ(0, g + 3, 55.0 / 96), B.PZ(6, g + 3, -5.0 / 32), B.PZ(4, g + 3, 5.0 / 8);
          B.PZ(3, g + 3, -5.0 / 32), B.PZ(2, g + 3, -15.0 / 32), B.PZ(1, g + 3, 5.0 / 16);
        } else {
          B.PZ(0, g + 4, 1.0 / 3);
        }
        if (c == 0) {
          const double s = 1.25;
          B.PZ(6, h + 3, -0.125 * s), B.PZ(4, h + 3, 0.25 * s), B.PZ(0, h + 3, -5.0 / 24 * s);
          B.PZ(1, h + 3, -0.125 * s), B.PZ(3, h + 3, 0.125 * s), B.PZ(5, h + 1, -0.125 * s);
          B.PZ(3, h + 1, 0.25 * s), B.PZ(0, h + 1, -5.0 / 24 * s), B.PZ(2, h + 1, -0.125 * s);
          B.PZ(4, h + 1, 0.125 * s);
        } else {
          B.PZ(0, h + 4, -1.0 / 3);
        }
      }

  // ---------------- TNS P_{T,jm} (redTime.cc:1163-1243) -------------------------------
  B.kpow = 0;
  B.scale = 1.0;
  B.row = ASM_PT0 + 0;  // j=2, m=2
  B.J(3, 4, 1.0 / 3), B.J(0, 4, -1.0 / 3);
  B.row = ASM_PT0 + 1;  // j=2, m=1
  B.J0(2, 7, 2 * -3.0 / 35, -2), B.J0(1, 7, 2 * 2.0 / 7, -2), B.J0(0, 7, 2 * -0.2, -2);
  B.row = ASM_PT0 + 2;  // j=2, m=0
  B.J0(6, 8, 5.0 / 231, -4), B.J0(5, 8, -9.0 / 77, -4), B.J0(4, 8, 5.0 / 21, -4), B.J0(3, 8, -1.0 / 7, -4);
  B.row = ASM_PT0 + 3;  // j=4, m=2
  B.J(3, 4, 1.0 / 3), B.J(2, 4, 2.0), B.J(0, 4, 5.0 / 3);
  B.row = ASM_PT0 + 4;  // j=4, m=1
  B.J(5, 5, -6.0 / 5), B.J(3, 7, 2.0), B.J(2, 5, 6.0 / 5), B.J(0, 7, -2.0);
  B.J0(2, 7, 12.0 / 7, -2), B.J0(1, 7, -40.0 / 7, -2), B.J0(0, 7, 4.0, -2);
  B.row = ASM_PT0 + 5;  // j=4, m=0
  B.J0(6, 8, -5.0 / 11, -4), B.J0(5, 8, 27.0 / 11, -4), B.J0(4, 8, -5.0, -4), B.J0(3, 8, 3.0, -4);
  B.J0(2, 8, -9.0 / 7, -2), B.J0(1, 8, 30.0 / 7, -2), B.J0(0, 8, -3.0, -2);
  B.J(6, 8, 27.0 / 70), B.J(3, 8, -9.0 / 7), B.J(0, 8, 9.0 / 10);
  B.row = ASM_PT0 + 6;  // j=6, m=1
  B.J0(2, 7, -2.0, -2), B.J0(1, 7, 20.0 / 3, -2), B.J0(0, 7, -14.0 / 3, -2);
  B.J(5, 5, 2.0), B.J(3, 7, -2.0 / 3), B.J(2, 7, 2.0), B.J(0, 7, 14.0 / 3);
  B.row = ASM_PT0 + 7;  // j=6, m=0
  B.J0(6, 8, 15.0 / 11, -4), B.J0(5, 8, -81.0 / 11, -4), B.J0(4, 8, 15.0, -4), B.J0(3, 8, -9.0, -4);
  B.J0(2, 8, 6.0, -2), B.J0(1, 8, -20.0, -2), B.J0(0, 8, 14.0, -2);
  B.J(6, 8, -39.0 / 35), B.J(5, 8, -6.0 / 5), B.J(3, 8, 47.0 / 7), B.J(2, 8, 6.0 / 5), B.J(0, 8, -28.0 / 5);
  B.row = ASM_PT0 + 8;  // j=8, m=0
  B.J0(6, 8, -1.0, -4), B.J0(5, 8, 27.0 / 5, -4), B.J0(4, 8, -11.0, -4), B.J0(3, 8, 33.0 / 5, -4);
  B.J0(2, 8, -27.0 / 5, -2), B.J0(1, 8, 18.0, -2), B.J0(0, 8, -63.0 / 5, -2);
  B.J(6, 8, 59.0 / 70), B.J(5, 8, 2.0), B.J(3, 8, -36.0 / 7), B.J(0, 8, 63.0 / 10);

  // ---------------- McDonald-Roy bias integrals (redTime.cc:1245-1278) ----------------
  B.row = ASM_PMR0 + 0, B.J(3, 0, 4.0 / 21), B.J(2, 0, 1.0), B.J(0, 0, 17.0 / 21);
  B.row = ASM_PMR0 + 1, B.J(3, 0, 8.0 / 21), B.J(2, 0, 1.0), B.J(0, 0, 13.0 / 21);
  B.row = ASM_PMR0 + 2;
  B.J(6, 0, 16.0 / 245), B.J(5, 0, 2.0 / 5), B.J(3, 0, 254.0 / 441), B.J(2, 0, 4.0 / 15), B.J(0, 0, 8.0 / 315);
  B.row = ASM_PMR0 + 3;
  B.J(6, 0, 32.0 / 245), B.J(5, 0, 2.0 / 5), B.J(3, 0, 214.0 / 441), B.J(2, 0, 4.0 / 15), B.J(0, 0, 16.0 / 315);
  B.row = ASM_PMR0 + 4, B.J(0, 0, 0.5), B.Jlo(-0.5);
  B.row = ASM_PMR0 + 5, B.J(3, 0, 1.0 / 3), B.Jlo(-1.0 / 3);
  B.row = ASM_PMR0 + 6, B.J(6, 0, 4.0 / 35), B.J(3, 0, 4.0 / 63), B.J(0, 0, 2.0 / 45), B.Jlo(-2.0 / 9);
  B.row = ASM_PMR0 + 7;
  B.PZ(6, 0, 0.5 * -15.0 / 128), B.PZ(4, 0, 0.5 * 15.0 / 32), B.PZ(3, 0, 0.5 * -15.0 / 128);
  B.PZ(2, 0, 0.5 * -45.0 / 128), B.PZ(1, 0, 0.5 * 15.0 / 64), B.PZ(0, 0, 0.5 * 55.0 / 128);
  return B.t;
}

}  // namespace

const std::vector<AsmTerm> &assembly_terms() {
  static const std::vector<AsmTerm> terms = build();
  return terms;
}

}  // namespace rtrg
