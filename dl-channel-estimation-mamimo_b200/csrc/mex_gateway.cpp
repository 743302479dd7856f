// MATLAB MEX gateway over the C ABI (include/mamimo.h).  Thin: argument checks + pointer hand-off.
//
//   mamimo_mex('create', cfg)                      cfg: struct with n_tx n_rx n_sc [n_ltf n_ps hidden d_out precision]
//   mamimo_mex('pilots', ltf_o, P)                 ltf_o = ltf(ind) [Nsc x 1] real, P [numSTS x nltf] real (helperGetP)
//   mamimo_mex('load', net, layer, W, b [, gamma, beta, mean, var])   W single [in x out]' (Keras kernel, row-major)
//   mamimo_mex('finalize')
//   hD = mamimo_mex('ls', rxData)                  rxData complex double [Nsc x nltf x Nr (x Npkt)]
//   [hD, Hr, Hi] = mamimo_mex('estimate', rxData)  Hr/Hi single [Nsc x Nt*Nr*Npkt] (column = pair row)
//   hM = mamimo_mex('lmmse', hD, tau, SNR)         LMMSE_ce over all pairs: tau = the `h` vector of LMMSE_ce (one per call)
//                                                  or a [Ntau x Npkt] matrix, SNR(i) in dB [Nr (x Npkt)]
//   mamimo_mex('ofdm', fftLen, cpLen, symOffset, carriers)   ofdmdemod parameters; carriers = prm.CarriersLocations
//   Y = mamimo_mex('demod', x)                     x complex double [nltf*(fftLen+cpLen) x Nr (x Npkt)] (inputRXSig) ->
//                                                  Y complex single [Nsc x nltf x Nr (x Npkt)] (= rxOFDM(:,1:nltf,:))
//   [Fbb, Frf, idx] = mamimo_mex('omphyb', hD, Ns, NtRF, At)   precoding-only omphybweights (pg/omphybweights.m:1,150-163) for
//                                                  all subcarriers: hD complex double [Nsc x Nt x Nr (x Npkt)], At [Nt x nRays] or
//                                                  the reference's AtExp [Nsc x Nt x nRays] (page 1 is used: the call site repeats
//                                                  one At, BER_test :366-369) -> Fbb [Nsc x Ns x NtRF], Frf [Nsc x NtRF x Nt],
//                                                  idx [Nsc x NtRF] 1-based dictionary columns (0 = loop stopped before)
//   ltf = mamimo_mex('ltf')                        256 x 1 tone table (no engine needed)
//   mamimo_mex('destroy')
//
// 'ls' replaces the loop body of pg/helperMIMOChannelEstimate.m:33-36; the .m shim of the original name is in
// matlab/helperMIMOChannelEstimate.m.  MATLAB's column-major [Nsc x nltf x Nr x Npkt] IS the engine's C-order
// [Npkt][Nr][nltf][Nsc], so buffers are handed over without a copy (interleaved complex, -R2018a).
// Errors follow the reference's error(message(...)) style through mexErrMsgIdAndTxt, which long-jumps:
// nothing heap-allocated is live at those points, and the engine persists in a static freed by mexAtExit.
#include <string.h>

#include "mex.h"
#include "mamimo.h"

static mamimo_engine* g_engine = NULL;

static void at_exit(void) {
  if (g_engine) { mamimo_destroy(g_engine); g_engine = NULL; }
}

static void fail(mamimo_status s) {
  if (s != MAMIMO_OK) mexErrMsgIdAndTxt("mamimo:engine", "%s: %s", mamimo_status_string(s), mamimo_last_error(g_engine));
}

static int field_int(const mxArray* st, const char* name, int dflt) {
  const mxArray* f = mxGetField(st, 0, name);
  return (f && !mxIsEmpty(f)) ? (int)mxGetScalar(f) : dflt;
}

static void need_engine(void) {
  if (!g_engine) mexErrMsgIdAndTxt("mamimo:state", "call mamimo_mex('create', cfg) first");
}

// rxData dims -> n_pkt; checks [Nsc x nltf x Nr (x Npkt)] against the engine configuration
static mwSize check_rx(const mxArray* rx, const mamimo_config* c) {
  if (!mxIsComplex(rx) || !mxIsDouble(rx)) mexErrMsgIdAndTxt("mamimo:type", "rxData must be complex double");
  const mwSize nd = mxGetNumberOfDimensions(rx);
  const mwSize* d = mxGetDimensions(rx);
  const mwSize nr = nd >= 3 ? d[2] : 1, np = nd >= 4 ? d[3] : 1;
  if (nd > 4 || d[0] != (mwSize)c->n_sc || d[1] != (mwSize)c->n_ltf || nr != (mwSize)c->n_rx)
    mexErrMsgIdAndTxt("mamimo:size", "rxData must be [Nsc x nltf x Nr (x Npkt)] = [%d x %d x %d]", c->n_sc, c->n_ltf, c->n_rx);
  return np;
}

static mamimo_config g_cfg;
static int g_fft = 0, g_cp = 0;

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
  if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("mamimo:usage", "first argument must be a command string");
  char cmd[32];
  {
    char* s = mxArrayToString(prhs[0]);
    strncpy(cmd, s ? s : "", sizeof(cmd) - 1);
    cmd[sizeof(cmd) - 1] = 0;
    mxFree(s);
  }
  if (!strcmp(cmd, "create")) {
    if (nrhs < 2 || !mxIsStruct(prhs[1])) mexErrMsgIdAndTxt("mamimo:usage", "create needs a cfg struct");
    at_exit();
    mamimo_config_init(&g_cfg);
    g_cfg.n_tx = field_int(prhs[1], "n_tx", 0);
    g_cfg.n_rx = field_int(prhs[1], "n_rx", 0);
    g_cfg.n_sc = field_int(prhs[1], "n_sc", 0);
    g_cfg.n_ltf = field_int(prhs[1], "n_ltf", g_cfg.n_tx);
    g_cfg.n_ps = field_int(prhs[1], "n_ps", 1);
    g_cfg.precision = field_int(prhs[1], "precision", MAMIMO_PREC_FP16X3);
    g_cfg.d_out = field_int(prhs[1], "d_out", 0);
    g_cfg.d_in = g_cfg.d_out > 0 ? g_cfg.n_sc : 0;
    const mxArray* h = mxGetField(prhs[1], 0, "hidden");
    if (h && mxIsDouble(h)) {
      g_cfg.n_hidden = (int)mxGetNumberOfElements(h);
      if (g_cfg.n_hidden > MAMIMO_MAX_HIDDEN) mexErrMsgIdAndTxt("mamimo:size", "too many hidden layers");
      for (int i = 0; i < g_cfg.n_hidden; ++i) g_cfg.hidden[i] = (int)mxGetDoubles(h)[i];
    }
    mamimo_status s = mamimo_create(&g_cfg, &g_engine);
    if (s != MAMIMO_OK) mexErrMsgIdAndTxt("mamimo:engine", "%s: %s", mamimo_status_string(s), mamimo_last_error(NULL));
    mexAtExit(at_exit);
  } else if (!strcmp(cmd, "destroy")) {
    at_exit();
  } else if (!strcmp(cmd, "ltf")) {               // 256 x 1 VHT-LTF tone table (helperMIMOChannelEstimate.m:16-23)
    int8_t t[256];
    mamimo_vht_ltf256(t);
    plhs[0] = mxCreateDoubleMatrix(256, 1, mxREAL);
    for (int i = 0; i < 256; ++i) mxGetDoubles(plhs[0])[i] = t[i];
  } else if (!strcmp(cmd, "pilots")) {
    need_engine();
    if (nrhs < 3 || !mxIsDouble(prhs[1]) || !mxIsDouble(prhs[2]) || mxIsComplex(prhs[1]) || mxIsComplex(prhs[2]))
      mexErrMsgIdAndTxt("mamimo:usage", "pilots needs real double ltf_o and P");
    const size_t np = mxGetNumberOfElements(prhs[1]), nP = mxGetNumberOfElements(prhs[2]);
    if (nP != (size_t)g_cfg.n_tx * g_cfg.n_ltf) mexErrMsgIdAndTxt("mamimo:size", "P must be [numSTS x nltf]");
    static double xp[2 * 65536], Pm[2 * 64 * 64];      // double tables: 'ls' on complex double runs in FP64 like MATLAB
    if (np > 65536) mexErrMsgIdAndTxt("mamimo:size", "too many pilot tones");
    for (size_t i = 0; i < np; ++i) { xp[2 * i] = mxGetDoubles(prhs[1])[i]; xp[2 * i + 1] = 0.0; }
    for (int j = 0; j < g_cfg.n_tx; ++j)            // MATLAB column-major P(j,n) -> row-major [tx][ltf]
      for (int n = 0; n < g_cfg.n_ltf; ++n) {
        Pm[2 * (j * g_cfg.n_ltf + n)] = mxGetDoubles(prhs[2])[n * g_cfg.n_tx + j];
        Pm[2 * (j * g_cfg.n_ltf + n) + 1] = 0.0;
      }
    fail(mamimo_set_pilots_f64(g_engine, xp, Pm));
  } else if (!strcmp(cmd, "load")) {
    need_engine();
    if (nrhs != 5 && nrhs != 9) mexErrMsgIdAndTxt("mamimo:usage", "load(net, layer, W, b [, gamma, beta, mean, var])");
    for (int i = 3; i < nrhs; ++i)
      if (!mxIsSingle(prhs[i])) mexErrMsgIdAndTxt("mamimo:type", "weights must be single");
    const float* bn[4] = {NULL, NULL, NULL, NULL};
    for (int i = 0; i < 4 && nrhs == 9; ++i) bn[i] = mxGetSingles(prhs[5 + i]);
    fail(mamimo_load_layer(g_engine, (int)mxGetScalar(prhs[1]), (int)mxGetScalar(prhs[2]), mxGetSingles(prhs[3]),
                           mxGetSingles(prhs[4]), bn[0], bn[1], bn[2], bn[3]));
  } else if (!strcmp(cmd, "finalize")) {
    need_engine();
    fail(mamimo_finalize_weights(g_engine));
  } else if (!strcmp(cmd, "ls") || !strcmp(cmd, "estimate")) {
    need_engine();
    if (nrhs < 2) mexErrMsgIdAndTxt("mamimo:usage", "rxData missing");
    const mwSize np = check_rx(prhs[1], &g_cfg);
    const mwSize dims[4] = {(mwSize)g_cfg.n_sc, (mwSize)g_cfg.n_tx, (mwSize)g_cfg.n_rx, np};
    plhs[0] = mxCreateNumericArray(np > 1 ? 4 : 3, dims, mxDOUBLE_CLASS, mxCOMPLEX);   // hD [Nsc x numSTS x Nr (x Npkt)]
    fail(mamimo_ls_estimate(g_engine, mxGetComplexDoubles(prhs[1]), MAMIMO_C128, MAMIMO_MEM_HOST, (int64_t)np,
                            mxGetComplexDoubles(plhs[0]), MAMIMO_C128, MAMIMO_MEM_HOST, NULL));
    if (!strcmp(cmd, "estimate")) {
      if (nlhs < 3) mexErrMsgIdAndTxt("mamimo:usage", "[hD, Hr, Hi] = mamimo_mex('estimate', rxData)");
      const mwSize od[2] = {(mwSize)g_cfg.d_out, (mwSize)g_cfg.n_tx * g_cfg.n_rx * np};   // column = pair row
      plhs[1] = mxCreateNumericArray(2, od, mxSINGLE_CLASS, mxREAL);
      plhs[2] = mxCreateNumericArray(2, od, mxSINGLE_CLASS, mxREAL);
      fail(mamimo_estimate(g_engine, mxGetComplexDoubles(prhs[1]), MAMIMO_C128, (int64_t)np, NULL,
                           mxGetSingles(plhs[1]), mxGetSingles(plhs[2]), MAMIMO_MEM_HOST, NULL));
    }
  } else if (!strcmp(cmd, "ofdm")) {             // ofdmdemod(x, FFT, CP, symOffset, null, pilot), pg/generate_maMIMO_LTF.m:336-338
    need_engine();
    if (nrhs < 5 || !mxIsDouble(prhs[4])) mexErrMsgIdAndTxt("mamimo:usage", "mamimo_mex('ofdm', fftLen, cpLen, symOffset, carriers)");
    if (mxGetNumberOfElements(prhs[4]) != (size_t)g_cfg.n_sc) mexErrMsgIdAndTxt("mamimo:size", "need Nsc carrier indices");
    static int32_t car[65536];
    for (int i = 0; i < g_cfg.n_sc; ++i) car[i] = (int32_t)mxGetDoubles(prhs[4])[i];
    g_fft = (int)mxGetScalar(prhs[1]); g_cp = (int)mxGetScalar(prhs[2]);
    fail(mamimo_set_ofdm(g_engine, g_fft, g_cp, (int)mxGetScalar(prhs[3]), car));
  } else if (!strcmp(cmd, "demod")) {
    need_engine();
    if (nrhs < 2 || !mxIsComplex(prhs[1]) || !mxIsDouble(prhs[1])) mexErrMsgIdAndTxt("mamimo:type", "x must be complex double");
    const mwSize nd = mxGetNumberOfDimensions(prhs[1]);
    const mwSize* d = mxGetDimensions(prhs[1]);
    const mwSize nr = nd >= 2 ? d[1] : 1, np = nd >= 3 ? d[2] : 1;
    if (g_fft == 0 || d[0] != (mwSize)g_cfg.n_ltf * (g_fft + g_cp) || nr != (mwSize)g_cfg.n_rx)
      mexErrMsgIdAndTxt("mamimo:size", "x must be [nltf*(fftLen+cpLen) x Nr (x Npkt)]; call mamimo_mex('ofdm', ...) first");
    const mwSize od[4] = {(mwSize)g_cfg.n_sc, (mwSize)g_cfg.n_ltf, (mwSize)g_cfg.n_rx, np};
    plhs[0] = mxCreateNumericArray(np > 1 ? 4 : 3, od, mxSINGLE_CLASS, mxCOMPLEX);
    fail(mamimo_ofdm_demod(g_engine, mxGetComplexDoubles(prhs[1]), MAMIMO_C128, (int64_t)np, mxGetComplexSingles(plhs[0]),
                           MAMIMO_MEM_HOST, NULL));
  } else if (!strcmp(cmd, "lmmse")) {            // isMMSE branch of pg/helperMIMOChannelEstimate.m:37-39
    need_engine();
    if (nrhs < 4) mexErrMsgIdAndTxt("mamimo:usage", "hM = mamimo_mex('lmmse', hD, tau, SNR)");
    const mxArray* hd = prhs[1];
    if (!mxIsComplex(hd) || !mxIsDouble(hd)) mexErrMsgIdAndTxt("mamimo:type", "hD must be complex double");
    const mwSize nd = mxGetNumberOfDimensions(hd);
    const mwSize* d = mxGetDimensions(hd);
    const mwSize nr = nd >= 3 ? d[2] : 1, np = nd >= 4 ? d[3] : 1;
    if (nd > 4 || d[0] != (mwSize)g_cfg.n_sc || d[1] != (mwSize)g_cfg.n_tx || nr != (mwSize)g_cfg.n_rx)
      mexErrMsgIdAndTxt("mamimo:size", "hD must be [Nsc x numSTS x Nr (x Npkt)]");
    if (!mxIsDouble(prhs[2]) || !mxIsDouble(prhs[3]) || mxIsComplex(prhs[3]))
      mexErrMsgIdAndTxt("mamimo:type", "tau must be double, SNR real double");
    const size_t n_tau_all = mxGetNumberOfElements(prhs[2]), n_snr = mxGetNumberOfElements(prhs[3]);
    if (n_snr != (size_t)nr && n_snr != (size_t)(nr * np)) mexErrMsgIdAndTxt("mamimo:size", "SNR must be [Nr] or [Nr x Npkt]");
    const bool tau_per_pkt = np > 1 && mxGetN(prhs[2]) == np && mxGetM(prhs[2]) > 1;
    const size_t n_tau = tau_per_pkt ? mxGetM(prhs[2]) : n_tau_all;
    static double t_rms[65536], snr[65536];
    if (np > 65536 || nr * np > 65536) mexErrMsgIdAndTxt("mamimo:size", "batch too large for one call");
    const int cplx = mxIsComplex(prhs[2]);
    const double* tp = cplx ? (const double*)mxGetComplexDoubles(prhs[2]) : mxGetDoubles(prhs[2]);
    for (mwSize p = 0; p < np; ++p)
      t_rms[p] = mamimo_tau_rms(tp + (tau_per_pkt ? p * n_tau * (cplx ? 2 : 1) : 0), (int32_t)n_tau, cplx);
    for (mwSize p = 0; p < np; ++p)                   // MATLAB [Nr x Npkt] column-major == [pkt][rx]
      for (mwSize i = 0; i < nr; ++i) snr[p * nr + i] = mxGetDoubles(prhs[3])[n_snr == (size_t)nr ? i : p * nr + i];
    plhs[0] = mxCreateNumericArray(nd, d, mxDOUBLE_CLASS, mxCOMPLEX);
    fail(mamimo_lmmse(g_engine, mxGetComplexDoubles(hd), MAMIMO_C128, (int64_t)np, t_rms, snr,
                      mxGetComplexDoubles(plhs[0]), MAMIMO_C128, MAMIMO_MEM_HOST, NULL));
  } else if (!strcmp(cmd, "omphyb")) {           // [Fbb,Frf] = omphybweights(Hchann_in,Ns,NtRF,At), pg/BER_test_maMIMO_LTF.m:372
    need_engine();
    if (nrhs < 5) mexErrMsgIdAndTxt("mamimo:usage", "[Fbb, Frf, idx] = mamimo_mex('omphyb', hD, Ns, NtRF, At)");
    const mxArray* hd = prhs[1];
    if (!mxIsComplex(hd) || !mxIsDouble(hd)) mexErrMsgIdAndTxt("mamimo:type", "hD must be complex double");
    const mwSize nd = mxGetNumberOfDimensions(hd);
    const mwSize* d = mxGetDimensions(hd);
    const mwSize nsc = (mwSize)g_cfg.n_sc, nt = (mwSize)g_cfg.n_tx;
    const mwSize nr = nd >= 3 ? d[2] : 1, np = nd >= 4 ? d[3] : 1;
    if (nd > 4 || d[0] != nsc || d[1] != nt || nr != (mwSize)g_cfg.n_rx)
      mexErrMsgIdAndTxt("mamimo:size", "hD must be [Nsc x Nt x Nr (x Npkt)] = [%d x %d x %d]", g_cfg.n_sc, g_cfg.n_tx, g_cfg.n_rx);
    const int ns = (int)mxGetScalar(prhs[2]), nrf = (int)mxGetScalar(prhs[3]);
    if (ns < 1 || ns > (int)nr || ns > nrf)      // omphybweights.m:116-117 (NS <= NTRF); rank(H) <= Nr
      mexErrMsgIdAndTxt("mamimo:size", "need 1 <= NS <= min(NTRF, Nr)");
    const mxArray* at = prhs[4];
    if (!mxIsComplex(at) || !mxIsDouble(at)) mexErrMsgIdAndTxt("mamimo:type", "At must be complex double");
    const mwSize and_ = mxGetNumberOfDimensions(at);
    const mwSize* ad = mxGetDimensions(at);
    mwSize n_rays = 0;
    mxArray* at2 = NULL;                          // [Nt x nRays]: column-major == the engine's [ray][tx] rows
    const mxComplexDouble* rows = NULL;
    if (and_ == 2 && ad[0] == nt) {
      n_rays = ad[1];
      rows = mxGetComplexDoubles(at);
    } else if (and_ == 3 && ad[0] == nsc && ad[1] == nt) {
      n_rays = ad[2];
      const mwSize dd[2] = {nt, n_rays};
      at2 = mxCreateNumericArray(2, dd, mxDOUBLE_CLASS, mxCOMPLEX);
      mxComplexDouble* w = mxGetComplexDoubles(at2);
      const mxComplexDouble* src = mxGetComplexDoubles(at);
      for (mwSize i = 0; i < nt * n_rays; ++i) w[i] = src[i * nsc];       // At(1, t, ray)
      rows = w;
    } else {
      mexErrMsgIdAndTxt("mamimo:size", "At must be [Nt x nRays] or [Nsc x Nt x nRays]");
    }
    if (nrf > (int)n_rays) mexErrMsgIdAndTxt("mamimo:size", "NTRF exceeds the number of dictionary columns");
    fail(mamimo_set_steering_dictionary(g_engine, (const double*)rows, (int32_t)n_rays));
    const mwSize sd[3] = {nsc, nr, np};
    mxArray* sig = mxCreateNumericArray(3, sd, mxDOUBLE_CLASS, mxREAL);
    mxArray* v1 = mxCreateNumericArray(nd, d, mxDOUBLE_CLASS, mxCOMPLEX);
    fail(mamimo_svd(g_engine, mxGetComplexDoubles(hd), MAMIMO_C128, (int64_t)np, mxGetDoubles(sig), mxGetComplexDoubles(v1),
                    MAMIMO_C128, MAMIMO_MEM_HOST, NULL));
    const mwSize id[3] = {nsc, (mwSize)nrf, np};
    mxArray* ix = mxCreateNumericArray(3, id, mxSINGLE_CLASS, mxREAL);    // 4-byte cells: int32 indices
    mxArray* er = mxCreateNumericArray(3, id, mxSINGLE_CLASS, mxREAL);
    const mwSize fd[4] = {nsc, (mwSize)nrf, (mwSize)ns, np};
    mxArray* fb = mxCreateNumericArray(4, fd, mxDOUBLE_CLASS, mxCOMPLEX);  // engine order [pkt][s][j][k]
    fail(mamimo_omp(g_engine, mxGetComplexDoubles(v1), MAMIMO_C128, (int32_t)nr, (int64_t)np, ns, nrf,
                    (int32_t*)mxGetSingles(ix), mxGetSingles(er), mxGetComplexDoubles(fb), MAMIMO_C128, MAMIMO_MEM_HOST, NULL));
    const int32_t* sel = (const int32_t*)mxGetSingles(ix);
    const mxComplexDouble* f = mxGetComplexDoubles(fb);
    const mwSize od0[4] = {nsc, (mwSize)ns, (mwSize)nrf, np};             // Fbb(k, s, j): omphybweights.m:154,196
    plhs[0] = mxCreateNumericArray(np > 1 ? 4 : 3, od0, mxDOUBLE_CLASS, mxCOMPLEX);
    mxComplexDouble* o0 = mxGetComplexDoubles(plhs[0]);
    for (mwSize p = 0; p < np; ++p)
      for (mwSize j = 0; j < (mwSize)nrf; ++j)
        for (mwSize q = 0; q < (mwSize)ns; ++q)
          memcpy(o0 + ((p * nrf + j) * ns + q) * nsc, f + ((p * ns + q) * nrf + j) * nsc, nsc * sizeof(mxComplexDouble));
    if (nlhs >= 2) {                                                       // Frf(k, j, t) = At(t, idx(k, j)): :155,197
      const mwSize od1[4] = {nsc, (mwSize)nrf, nt, np};
      plhs[1] = mxCreateNumericArray(np > 1 ? 4 : 3, od1, mxDOUBLE_CLASS, mxCOMPLEX);
      mxComplexDouble* o1 = mxGetComplexDoubles(plhs[1]);
      for (mwSize p = 0; p < np; ++p)
        for (mwSize t = 0; t < nt; ++t)
          for (mwSize j = 0; j < (mwSize)nrf; ++j)
            for (mwSize k = 0; k < nsc; ++k) {
              const int32_t c = sel[(p * nrf + j) * nsc + k];
              if (c >= 0) o1[((p * nt + t) * nrf + j) * nsc + k] = rows[(mwSize)c * nt + t];
            }
    }
    if (nlhs >= 3) {
      plhs[2] = mxCreateNumericArray(np > 1 ? 3 : 2, id, mxDOUBLE_CLASS, mxREAL);
      double* o2 = mxGetDoubles(plhs[2]);
      for (mwSize i = 0; i < nsc * nrf * np; ++i) o2[i] = (double)(sel[i] + 1);
    }
    mxDestroyArray(sig); mxDestroyArray(v1); mxDestroyArray(ix); mxDestroyArray(er); mxDestroyArray(fb);
    if (at2) mxDestroyArray(at2);
  } else {
    mexErrMsgIdAndTxt("mamimo:usage", "unknown command '%s'", cmd);
  }
}
