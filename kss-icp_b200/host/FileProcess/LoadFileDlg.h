// FileProcess/LoadFileDlg.h -- empty stand-in for the Win32 file dialog (windows.h / commdlg.h), unused on the path.
#pragma once
