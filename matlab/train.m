function model = train(model,X,Y,varargin)
% Drop-in for GPz/train.m:1-81.  The whole optimisation -- minFunc's L-BFGS with the Wolfe line search as train.m:42-48
% configures it, and the best-theta / maxAttempts rule of GPz/callBack.m -- runs on the GPU inside one MEX call
% (gpz_train); theta never travels during the run.  Afterwards w, iSigma_w and the priors of the last and the best
% theta are re-fitted as train.m:53-79 does.
n = size(X,1);
pnames = {'maxIter' 'maxAttempts' 'omega' 'training' 'validation' 'Psi' 'display'};
defaults = {200 inf ones(n,1) true(n,1) [] [] true};
[maxIter,maxAttempts,omega,training,validation,Psi,display] = internal.stats.parseArgs(pnames,defaults,varargin{:});
m = model.m; k = model.k; d = size(X,2);
Y = bsxfun(@minus,Y,model.muY);
X = bsxfun(@rdivide,bsxfun(@minus,X,model.muX),model.sdX);
if ~isempty(Psi), Psi = fixPsi(Psi,n,model.sdX,model.method); end
h = gpz_b200_mex('create',model,X,Y,Psi,omega,training,validation);
cleanup = onCleanup(@() gpz_b200_mex('destroy',h));
[theta,best_theta,best_valid,info] = gpz_b200_mex('train',h,model.last.theta,model.best.theta,model.best.LL, ...
    maxIter,maxAttempts,isempty(validation),display);
sets = {'last','best'}; thetas = {theta,best_theta};
for s = 1:2
    th = thetas{s};
    [~,w,iSigma_w] = gpz_b200_mex('fit',h,th,model);
    r = struct('theta',th,'w',w,'iSigma_w',iSigma_w,'priors',ones(1,m)/m,'P',reshape(th(1:m*d),m,d));
    if model.heteroscedastic
        o = m*d+model.g_dim+m*k+k;
        r.v = reshape(th(o+1:o+m*k),m,k);
    end
    if s == 2, r.LL = model.best.LL; end          % train.m never writes best.LL back: every call restarts from init's -inf
    model.(sets{s}) = r;                          % stored before the priors: a failing EM must not cost the training result
    try
        model.(sets{s}).priors = gpz_b200_mex('get_prior',h,th,model);   % train.m:59,74
    catch err
        warning('gpz_b200:prior','getPrior failed (%s); priors left uniform',err.message);
    end
end
model.train_info = [info best_valid];
end
