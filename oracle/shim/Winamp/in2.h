/* stand-in for the Winamp SDK input-plugin header: the DSP path only passes In_Module pointers around (typedef in windows.h shim). Test infrastructure only. */
