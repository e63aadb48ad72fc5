/* intentionally empty: Win32 header not needed by the DSP path (oracle shim, test infrastructure only) */
