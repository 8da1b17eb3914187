function [nlogML,grad,w,iSigma_w,PHI] = GPz(theta,model,X,Y,Psi,omega,training,validation)
% Drop-in replacement for GPz/GPz.m:1 -- same signature, same nargout switch (GPz.m:84-87), same four
% global side-channel statistics (GPz.m:3-7,236-259) so that GPz/callBack.m works unmodified.
% The data arguments are constant for a whole optimisation (train.m:40), so the device context is
% cached across calls and only theta travels; `clear GPz` (or a different data set) rebuilds it.
global trainRMSE trainLL validRMSE validLL
persistent h key
newkey = [size(X) size(Y) numel(Psi) sum(training(:)) sum(validation(:)) X(1) Y(1) X(end) Y(end)];
if isempty(h) || ~isequal(key,newkey)
    if ~isempty(h), gpz_b200_mex('destroy',h); end
    h = gpz_b200_mex('create',model,X,Y,Psi,omega,training,validation);
    key = newkey;
end
if isempty(Y)                                   % GPz.m:34-40
    nlogML = 0; grad = 0; w = 0; iSigma_w = 0; return
end
if nargout > 2                                  % fit exit, GPz.m:84-87
    [nlogML,w,iSigma_w] = gpz_b200_mex('fit',h,theta,model);
    grad = 0;
    if nargout > 4, PHI = gpz_b200_mex('phi',h,theta,0,model); end
else
    [nlogML,grad,stats] = gpz_b200_mex('eval',h,theta);
    trainRMSE = stats(1); trainLL = stats(2);
    if ~isempty(validation), validRMSE = stats(3); validLL = stats(4); end
end
end
