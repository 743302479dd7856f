function [hD,P,ltf_o,hDmmse] = helperMIMOChannelEstimate(rxData,prm,Nps,tau,SNR,isMMSE)
% Drop-in shim with the reference's signature (pg/helperMIMOChannelEstimate.m:1) that routes the
% LS estimate through the B200 engine (mamimo_mex).  rxData may also be [Nsc x nltf x Nr x Npkt]
% when the call is hoisted out of the per-packet loop of generate_maMIMO_LTF.m:197,342.
% P and the LTF tone table stay MATLAB-side inputs (helperGetP is a MathWorks example helper).
persistent cfgKey
numSTS = prm.numSTS;
P = helperGetP(numSTS);
ltf = mamimo_mex('ltf');                   % same 256-tone table as the reference (:16-23)
ind = prm.CarriersLocations;
ltf_o = ltf(ind);
key = [numSTS size(rxData,3) numel(ind)];
if isempty(cfgKey) || ~isequal(cfgKey,key)
    mamimo_mex('create', struct('n_tx',numSTS,'n_rx',size(rxData,3),'n_sc',numel(ind),'n_ltf',size(rxData,2)));
    cfgKey = key;
end
mamimo_mex('pilots', double(ltf_o), double(P(1:numSTS,1:numSTS)));
hD = mamimo_mex('ls', complex(double(rxData)));
hDmmse = complex(zeros(size(hD)));
if isMMSE                                   % LMMSE_ce for every pair (:37-39), one FP64 solve per (packet, rx)
    if Nps ~= 1, error('mamimo:unsupported','create the engine with n_ps = Nps for comb pilots'); end
    hDmmse = mamimo_mex('lmmse', hD, double(tau), double(SNR));
end
end
