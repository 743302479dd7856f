/* Minimal stand-in for MATLAB's mex.h (R2018a interleaved-complex API), ONLY so that
 * mex_gateway.cpp can be syntax-checked in an image without MATLAB.  Declarations follow the
 * documented MATLAB C Matrix API; nothing here is linked or shipped.  A real build uses
 *   mex -R2018a mex_gateway.cpp -I<repo>/include -L<pkg> -lmamimo_b200
 * against MathWorks' own header. */
#ifndef MAMIMO_MEX_STUB_H_
#define MAMIMO_MEX_STUB_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef struct { double real, imag; } mxComplexDouble;
typedef struct { float real, imag; } mxComplexSingle;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7 } mxClassID;
mwSize mxGetNumberOfDimensions(const mxArray*);
const mwSize* mxGetDimensions(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
int mxIsComplex(const mxArray*);
int mxIsDouble(const mxArray*);
int mxIsSingle(const mxArray*);
int mxIsChar(const mxArray*);
int mxIsStruct(const mxArray*);
int mxIsEmpty(const mxArray*);
double mxGetScalar(const mxArray*);
char* mxArrayToString(const mxArray*);
void mxFree(void*);
mxArray* mxGetField(const mxArray*, mwSize, const char*);
mxComplexDouble* mxGetComplexDoubles(const mxArray*);
mxComplexSingle* mxGetComplexSingles(const mxArray*);
double* mxGetDoubles(const mxArray*);
float* mxGetSingles(const mxArray*);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexAtExit(void (*)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
