function [PHI,Gamma,lnBeta_i,N] = getPHI(X,Psi,theta,model,selection)
% Drop-in for GPz/getPHI.m:1 (PHI and lnBeta_i on the GPU; Gamma unpacked on the host as getPHI.m:26-40).
if isempty(selection), selection = true(size(X,1),1); end
h = gpz_b200_mex('create',model,X,zeros(size(X,1),model.k),Psi,[],selection,[]);
if nargout > 3
    [PHI,lnBeta_i,N] = gpz_b200_mex('phi',h,theta,0,model);
else
    [PHI,lnBeta_i] = gpz_b200_mex('phi',h,theta,0,model);
end
gpz_b200_mex('destroy',h);
m = model.m; d = model.d;
switch model.method
    case 'GL', Gamma = repmat(theta(m*d+1),m,d);
    case 'VL', Gamma = repmat(theta(m*d+1:m*d+m),1,d);
    case 'GD', Gamma = repmat(theta(m*d+1:m*d+d)',m,1);
    case 'VD', Gamma = reshape(theta(m*d+1:m*d+m*d),m,d);
    case 'GC', Gamma = reshape(repmat(reshape(theta(m*d+1:m*d+d*d),d,d),1,m),d,d,m);
    case 'VC', Gamma = reshape(theta(m*d+1:m*d+d*d*m),d,d,m);
end
end
