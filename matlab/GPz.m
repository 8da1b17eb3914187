function [nlogML,grad,w,iSigma_w,PHI] = GPz(theta,model,X,Y,Psi,omega,training,validation)
% Drop-in replacement for GPz/GPz.m:1 -- same signature, same nargout switch (GPz.m:84-87), same four
% global side-channel statistics (GPz.m:3-7,236-259) so that GPz/callBack.m works unmodified.
% The data arguments are constant for a whole optimisation (train.m:40), so the device context is
% cached across calls and only theta travels.  The cache is keyed on EVERYTHING the context depends on:
% the model fields that define theta's layout, both masks in full, omega, and position-weighted
% checksums of X, Y and Psi -- a call with another m / method / omega / split / data set rebuilds the
% context instead of reusing a stale one (the MEX gateway additionally refuses a theta whose length does
% not match the context's model).  `clear GPz` drops the cache.
global trainRMSE trainLL validRMSE validLL
persistent h key
if isempty(Y)                                   % GPz.m:34-40
    nlogML = 0; grad = 0; w = 0; iSigma_w = 0; return
end
newkey = {model.d, model.k, model.m, model.method, logical(model.heteroscedastic), size(X), size(Y), size(Psi), ...
          logical(training(:)), logical(validation(:)), checksum(omega), checksum(X), checksum(Y), checksum(Psi)};
if isempty(h) || ~isequaln(key,newkey)
    if ~isempty(h), gpz_b200_mex('destroy',h); h = []; end
    h = gpz_b200_mex('create',model,X,Y,Psi,omega,training,validation);
    key = newkey;
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

function c = checksum(A)
% three moments that change under any edit, permutation or rescaling of A that matters: plain sum, sum
% weighted by position, sum of squares (NaN = missing value: counted separately, isequaln compares NaN == NaN)
if isempty(A), c = []; return; end
a = double(A(:));
miss = isnan(a);
a(miss) = 0;
pos = (1:numel(a))';
c = [sum(a), sum(a.*pos)/numel(a), sum(a.*a), sum(miss), sum(pos(miss))];
end
