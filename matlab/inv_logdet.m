function [Xi,logdet] = inv_logdet(X)
% Drop-in for GPz/inv_logdet.m:1 for symmetric positive definite X (blocked Cholesky on the GPU; NaN out
% if X is not positive definite, where the reference would return a truncated pseudo-inverse).
[Xi,logdet] = gpz_b200_mex('inv_logdet',X);
end
