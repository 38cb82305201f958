// Cells per tile (= cells per CTA pass) of the cell kernels, per degree.
#pragma once
namespace bp5 {
// cells per tile for each degree: fills the CTA's warps ((p+1)^2 threads per
// cell) while keeping >= 2-4 CTAs resident per SM.
// (-DBP5_CPT_Pn=... overrides one entry for tuning builds, scripts/tune_cpt.sh)
#ifndef BP5_CPT_P1
#define BP5_CPT_P1 32
#endif
#ifndef BP5_CPT_P2
#define BP5_CPT_P2 14
#endif
#ifndef BP5_CPT_P3
#define BP5_CPT_P3 8
#endif
#ifndef BP5_CPT_P4
#define BP5_CPT_P4 5
#endif
#ifndef BP5_CPT_P5
#define BP5_CPT_P5 3
#endif
#ifndef BP5_CPT_P6
#define BP5_CPT_P6 3
#endif
#ifndef BP5_CPT_P7
#define BP5_CPT_P7 2
#endif
#ifndef BP5_CPT_P8
#define BP5_CPT_P8 1
#endif
template <int P> struct TileCells;
template <> struct TileCells<1> { static constexpr int value = BP5_CPT_P1; };
template <> struct TileCells<2> { static constexpr int value = BP5_CPT_P2; };
template <> struct TileCells<3> { static constexpr int value = BP5_CPT_P3; };
template <> struct TileCells<4> { static constexpr int value = BP5_CPT_P4; };
template <> struct TileCells<5> { static constexpr int value = BP5_CPT_P5; };
template <> struct TileCells<6> { static constexpr int value = BP5_CPT_P6; };
template <> struct TileCells<7> { static constexpr int value = BP5_CPT_P7; };
template <> struct TileCells<8> { static constexpr int value = BP5_CPT_P8; };

// the on-the-fly-geometry kernel keeps 12-15 work arrays per cell in shared memory: fewer cells per tile
template <int P> struct OtfTileCells { static constexpr int value = TileCells<P>::value; };
template <> struct OtfTileCells<6> { static constexpr int value = 2; };
template <> struct OtfTileCells<7> { static constexpr int value = 2; };
}  // namespace bp5
