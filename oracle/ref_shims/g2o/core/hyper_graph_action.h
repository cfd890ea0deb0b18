/* shim: unused by the hot path */
