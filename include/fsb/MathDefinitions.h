// Host-side scalar definitions of the B200 drop-in (reference: include/MathDefinitions.h:7-19).
// Only what the simulation path uses: MyFloat and CLAMP.  smoothstep/gaussian are
// renderer helpers and out of scope.
#ifndef FSB_MATH_DEFINITIONS_H
#define FSB_MATH_DEFINITIONS_H

typedef float MyFloat; // the reference's USE_DOUBLE_PRECISION switch is commented out upstream

// The reference clamps THROUGH MyFloat (ints are converted to float and back,
// include/MathDefinitions.h:16-19); keep that contract for callers that rely on it.
inline MyFloat CLAMP(MyFloat value, MyFloat low, MyFloat high)
{
  return value < low ? low : (value > high ? high : value);
}

#endif
