function [mu,sigma,nu,beta_i,gamma,PHI,w,iSigma_w] = predict(X,model,varargin)
% Drop-in for GPz/predict.m:1-75: host-side selection / z-scoring / fixPsi as in the reference, the
% per-row work on the GPU: predictFull / predictNoisy / predictMissing / predictNoisyMissing for the diagonal and the
% covariance methods (predictDiag.m:58-295, predictCov.m:53-336).  Rows with missing values need set.priors (getPrior.m).
pnames = {'whichSet' 'Psi' 'selection'};
defaults = {'best' [] true(size(X,1),1)};
[whichSet,Psi,selection] = internal.stats.parseArgs(pnames,defaults,varargin{:});
if strcmp(whichSet,'best'), set = model.best; else, set = model.last; end
n = sum(selection);
X = X(selection,:);
if ~isempty(Psi)
    if model.method(2)=='C', Psi = Psi(:,:,selection); else, Psi = Psi(selection,:); end
end
X = bsxfun(@rdivide,bsxfun(@minus,X,model.muX),model.sdX);
Psi = fixPsi(Psi,n,model.sdX,model.method);
w = set.w; iSigma_w = set.iSigma_w;
[mu,nu,beta_i,gamma,PHI] = gpz_b200_mex('predict',model,set.theta,w,iSigma_w,X,Psi,set.priors);
sigma = nu+beta_i+gamma;
mu = bsxfun(@plus,mu,model.muY);
end
