function D = Dxy(X,Y)
% Drop-in for GPz/Dxy.m:1.
D = gpz_b200_mex('dxy',X,Y);
end
