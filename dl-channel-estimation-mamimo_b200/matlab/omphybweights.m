function [Fbb,Frf] = omphybweights(Hchann_in,Ns,NtRF,At)
% Drop-in shim with the reference's precoding-only signature (pg/omphybweights.m:1, called at
% pg/BER_test_maMIMO_LTF.m:372 as omphybweights(hDp{1},numSTS,numSTS,AtExp)) that runs the per-subcarrier SVD and
% the orthogonal matching pursuit of all subcarriers in one call on the B200 engine (mamimo_mex 'omphyb').
% Hchann_in [L x Nt x Nr] complex, At [L x Nt x nRays] (one dictionary repeated per subcarrier, :366-369) or
% [Nt x nRays].  Returns Fbb [L x Ns x NtRF], Frf [L x NtRF x Nt] like the reference (:154-155).
% The combiner outputs (Wbb, Wrf; 6-7 input arguments) stay with the MathWorks implementation.
narginchk(4,4);
[L,Nt,Nr] = size(Hchann_in);
H = complex(double(Hchann_in));
A = complex(double(At));
try
    [Fbb,Frf] = mamimo_mex('omphyb', H, Ns, NtRF, A);
catch err                                   % no engine yet, or one created for other dimensions (the estimator shim)
    if ~any(strcmp(err.identifier, {'mamimo:state','mamimo:size'})), rethrow(err); end
    mamimo_mex('create', struct('n_tx',Nt,'n_rx',Nr,'n_sc',L,'n_ltf',Nt));
    [Fbb,Frf] = mamimo_mex('omphyb', H, Ns, NtRF, A);
end
end
