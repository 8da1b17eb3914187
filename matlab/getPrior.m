function prior = getPrior(X,Sx,theta,model,set)
% Drop-in for GPz/getPrior.m:1 (EM over the normalised basis densities, on the GPU).
h = gpz_b200_mex('create',model,X,zeros(size(X,1),model.k),Sx,[],set,[]);
prior = gpz_b200_mex('get_prior',h,theta,model);
gpz_b200_mex('destroy',h);
end
