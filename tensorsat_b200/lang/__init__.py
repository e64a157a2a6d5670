"""Box languages.  ``fin_rel`` holds the device-resident relation/counting boxes."""
