"""unit conversions used on the hot path (values of lime/units.py:2-9)"""
au2fs = 2.41888432651e-2
au2k = 315775.13
au2ev = 27.211386
au2mev = 27211.386
au2wavenumber = 219474.6305
