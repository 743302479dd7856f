// Mini MEX runtime (see mex.h): data model + the error / at-exit behaviour mex_gateway.cpp relies on.
#include "mex.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <utility>
#include <vector>

#include "mex_runtime.hpp"

struct mxArray_tag {
  mxClassID cls = mxUNKNOWN_CLASS;
  bool cplx = false;
  std::vector<mwSize> dims;
  std::vector<unsigned char> data;                                  // column-major, complex interleaved
  std::vector<std::pair<std::string, mxArray*>> fields;             // 1x1 struct
};

namespace {
size_t elem_bytes(const mxArray* a) {
  const size_t b = a->cls == mxDOUBLE_CLASS ? 8 : a->cls == mxSINGLE_CLASS ? 4 : a->cls == mxCHAR_CLASS ? 2 : 0;
  return a->cplx ? 2 * b : b;
}
std::vector<void (*)(void)> g_atexit;
}  // namespace

namespace minimex {
void run_atexit() {
  std::vector<void (*)(void)> h;
  h.swap(g_atexit);
  for (auto it = h.rbegin(); it != h.rend(); ++it) (*it)();
}
size_t n_atexit() { return g_atexit.size(); }
}  // namespace minimex

extern "C" {

mwSize mxGetNumberOfDimensions(const mxArray* a) { return a->dims.size(); }
const mwSize* mxGetDimensions(const mxArray* a) { return a->dims.data(); }
size_t mxGetNumberOfElements(const mxArray* a) {
  size_t n = 1;
  for (mwSize d : a->dims) n *= d;
  return n;
}
size_t mxGetM(const mxArray* a) { return a->dims[0]; }
size_t mxGetN(const mxArray* a) {                                   // product of dims 2..end, as MATLAB defines it
  size_t n = 1;
  for (size_t i = 1; i < a->dims.size(); ++i) n *= a->dims[i];
  return n;
}
int mxIsComplex(const mxArray* a) { return a->cplx; }
int mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
int mxIsSingle(const mxArray* a) { return a->cls == mxSINGLE_CLASS; }
int mxIsChar(const mxArray* a) { return a->cls == mxCHAR_CLASS; }
int mxIsStruct(const mxArray* a) { return a->cls == mxSTRUCT_CLASS; }
int mxIsEmpty(const mxArray* a) { return mxGetNumberOfElements(a) == 0; }
double mxGetScalar(const mxArray* a) {                              // first (real) element, converted to double
  if (mxIsEmpty(a)) mexErrMsgIdAndTxt("MATLAB:mxGetScalar", "empty array");
  if (a->cls == mxDOUBLE_CLASS) return *reinterpret_cast<const double*>(a->data.data());
  if (a->cls == mxSINGLE_CLASS) return *reinterpret_cast<const float*>(a->data.data());
  if (a->cls == mxCHAR_CLASS) return *reinterpret_cast<const unsigned short*>(a->data.data());
  mexErrMsgIdAndTxt("MATLAB:mxGetScalar", "not a numeric array");
  return 0;
}
char* mxArrayToString(const mxArray* a) {
  if (a->cls != mxCHAR_CLASS) return nullptr;
  const size_t n = mxGetNumberOfElements(a);
  char* s = static_cast<char*>(malloc(n + 1));
  for (size_t i = 0; i < n; ++i) s[i] = static_cast<char>(reinterpret_cast<const unsigned short*>(a->data.data())[i]);
  s[n] = 0;
  return s;
}
void mxFree(void* p) { free(p); }
mxArray* mxGetField(const mxArray* a, mwIndex idx, const char* name) {
  if (a->cls != mxSTRUCT_CLASS || idx != 0) return nullptr;
  for (auto& f : a->fields)
    if (f.first == name) return f.second;
  return nullptr;
}
// typed accessors are strict like -R2018a's: wrong class / complexity is an error, not a reinterpretation
mxComplexDouble* mxGetComplexDoubles(const mxArray* a) {
  if (a->cls != mxDOUBLE_CLASS || !a->cplx) mexErrMsgIdAndTxt("MATLAB:mxGetComplexDoubles", "array is not complex double");
  return reinterpret_cast<mxComplexDouble*>(const_cast<unsigned char*>(a->data.data()));
}
mxComplexSingle* mxGetComplexSingles(const mxArray* a) {
  if (a->cls != mxSINGLE_CLASS || !a->cplx) mexErrMsgIdAndTxt("MATLAB:mxGetComplexSingles", "array is not complex single");
  return reinterpret_cast<mxComplexSingle*>(const_cast<unsigned char*>(a->data.data()));
}
double* mxGetDoubles(const mxArray* a) {
  if (a->cls != mxDOUBLE_CLASS || a->cplx) mexErrMsgIdAndTxt("MATLAB:mxGetDoubles", "array is not real double");
  return reinterpret_cast<double*>(const_cast<unsigned char*>(a->data.data()));
}
float* mxGetSingles(const mxArray* a) {
  if (a->cls != mxSINGLE_CLASS || a->cplx) mexErrMsgIdAndTxt("MATLAB:mxGetSingles", "array is not real single");
  return reinterpret_cast<float*>(const_cast<unsigned char*>(a->data.data()));
}
mxArray* mxCreateNumericArray(mwSize nd, const mwSize* dims, mxClassID cls, mxComplexity c) {
  mxArray* a = new mxArray_tag();
  a->cls = cls;
  a->cplx = c == mxCOMPLEX;
  a->dims.assign(dims, dims + nd);
  while (a->dims.size() < 2) a->dims.push_back(1);
  while (a->dims.size() > 2 && a->dims.back() == 1) a->dims.pop_back();   // MATLAB drops trailing singletons
  a->data.assign(mxGetNumberOfElements(a) * elem_bytes(a), 0);             // zero-initialised like MATLAB
  return a;
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
  const mwSize d[2] = {m, n};
  return mxCreateNumericArray(2, d, mxDOUBLE_CLASS, c);
}
mxArray* mxCreateString(const char* s) {
  const mwSize d[2] = {1, strlen(s)};
  mxArray* a = mxCreateNumericArray(2, d, mxCHAR_CLASS, mxREAL);
  for (size_t i = 0; i < d[1]; ++i) reinterpret_cast<unsigned short*>(a->data.data())[i] = static_cast<unsigned char>(s[i]);
  return a;
}
mxArray* mxCreateStructMatrix(mwSize m, mwSize n, int nfields, const char** names) {
  mxArray* a = new mxArray_tag();
  a->cls = mxSTRUCT_CLASS;
  a->dims = {m, n};
  for (int i = 0; i < nfields; ++i) a->fields.emplace_back(names[i], nullptr);
  return a;
}
void mxSetField(mxArray* a, mwIndex idx, const char* name, mxArray* v) {
  if (a->cls != mxSTRUCT_CLASS || idx != 0) return;
  for (auto& f : a->fields)
    if (f.first == name) { f.second = v; return; }
  a->fields.emplace_back(name, v);
}
void mxDestroyArray(mxArray* a) {
  if (!a) return;
  for (auto& f : a->fields) mxDestroyArray(f.second);
  delete a;
}
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw minimex::MexError{id ? id : "", buf};                       // control never returns to the caller, as in MATLAB
}
int mexAtExit(void (*fn)(void)) {
  for (auto h : g_atexit)
    if (h == fn) return 0;                                          // registering twice keeps one handler
  g_atexit.push_back(fn);
  return 0;
}

}  // extern "C"
