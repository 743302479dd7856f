/* Mini MEX runtime: the subset of MATLAB's C Matrix / MEX API (R2018a interleaved-complex) that
 * mex_gateway.cpp uses, IMPLEMENTED (mex_runtime.cpp) so the gateway's mexFunction can be EXECUTED in an image without
 * MATLAB: mxArray with class, complexity, column-major dims, interleaved complex data, char arrays, 1x1 structs with
 * named fields; mexErrMsgIdAndTxt leaves mexFunction the way MATLAB's long jump does (a C++ exception caught by the
 * harness, mex_harness.cpp); mexAtExit handlers run when the harness "clears" the MEX file.
 * Declarations follow the documented API so the same gateway source builds unchanged with
 *   mex -R2018a mex_gateway.cpp -I<repo>/include -L<pkg> -lmamimo_b200
 * against MathWorks' own header.  Test infrastructure: nothing here ships in libmamimo_b200.so. */
#ifndef MAMIMO_MINI_MEX_H_
#define MAMIMO_MINI_MEX_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef size_t mwIndex;
typedef struct { double real, imag; } mxComplexDouble;
typedef struct { float real, imag; } mxComplexSingle;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxUNKNOWN_CLASS = 0, mxSTRUCT_CLASS = 2, mxCHAR_CLASS = 4, mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7 } mxClassID;

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
mxArray* mxGetField(const mxArray*, mwIndex, const char*);
mxComplexDouble* mxGetComplexDoubles(const mxArray*);
mxComplexSingle* mxGetComplexSingles(const mxArray*);
double* mxGetDoubles(const mxArray*);
float* mxGetSingles(const mxArray*);
mxArray* mxCreateNumericArray(mwSize, const mwSize*, mxClassID, mxComplexity);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateString(const char*);
mxArray* mxCreateStructMatrix(mwSize, mwSize, int, const char**);
void mxSetField(mxArray*, mwIndex, const char*, mxArray*);
void mxDestroyArray(mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexAtExit(void (*)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
