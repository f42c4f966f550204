"""On-disk event format of the reference (row f2 of SURVEY.md 8f): HDF5 tables of compound records
(`WaveformPairs{coord i4[3], waveform i2[2*ns], ...}` with an `nevents` attribute), read without h5py / libhdf5."""
